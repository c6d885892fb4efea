from . import caffe_pb2  # noqa: F401
