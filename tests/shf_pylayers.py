"""Python layers for tests/test_gpu_pycaffe_protocol.py -- the layer classes of the reference's own protocol test
(caffe/python/caffe/test/test_python_layer.py:9-57), re-expressed against this repo's `caffe` drop-in."""
import caffe


class SimpleLayer(caffe.Layer):
    """A layer that just multiplies by ten"""

    def setup(self, bottom, top):
        pass

    def reshape(self, bottom, top):
        top[0].reshape(*bottom[0].data.shape)

    def forward(self, bottom, top):
        top[0].data[...] = 10 * bottom[0].data


class ExceptionLayer(caffe.Layer):
    def setup(self, bottom, top):
        raise RuntimeError


class ParameterLayer(caffe.Layer):
    def setup(self, bottom, top):
        self.blobs.add_blob(1)
        self.blobs[0].data[0] = 0

    def reshape(self, bottom, top):
        top[0].reshape(*bottom[0].data.shape)

    def forward(self, bottom, top):
        pass


class PhaseLayer(caffe.Layer):
    def setup(self, bottom, top):
        pass

    def reshape(self, bottom, top):
        top[0].reshape()

    def forward(self, bottom, top):
        top[0].data[()] = self.phase


class ParamStrLayer(caffe.Layer):
    """Adds the number in param_str (set before setup, python_layer.hpp:24-25)."""

    def setup(self, bottom, top):
        self.value = float(self.param_str)

    def reshape(self, bottom, top):
        top[0].reshape(*bottom[0].data.shape)

    def forward(self, bottom, top):
        top[0].data[...] = bottom[0].data + self.value
