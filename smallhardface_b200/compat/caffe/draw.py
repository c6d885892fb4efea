"""`caffe.draw` stand-in: ``lib/prototxt/manipulate.py:74,85`` calls ``draw_net_to_file`` to render the net with
pydot/graphviz, neither of which exists in this image.  A text listing of the layers is written instead so the
caller's file-exists expectations hold."""


def draw_net_to_file(caffe_net, filename, rankdir="LR", phase=None):
    with open(filename, "w") as f:
        f.write("# graph rendering unavailable (no graphviz); layers of net %r\n" % getattr(caffe_net, "name", ""))
        for l in caffe_net.layer:
            f.write("%s (%s): %s -> %s\n" % (l.name, l.type, list(l.bottom), list(l.top)))
