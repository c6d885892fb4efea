"""`caffe.proto.caffe_pb2` without protoc: real google.protobuf message classes built at import time from the
restated schema in smallhardface_b200/caffe_proto.py (field numbers of caffe/src/caffe/proto/caffe.proto), so
``text_format.Merge``, ``str(pb)``, repeated-field append/extend, ``ClearField`` -- everything
``lib/prototxt/manipulate.py`` does -- works on them."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from smallhardface_b200.caffe_proto import DEFAULTS, ENUMS, SCHEMA

_T = descriptor_pb2.FieldDescriptorProto
_SCALAR = {"int32": _T.TYPE_INT32, "int64": _T.TYPE_INT64, "uint32": _T.TYPE_UINT32, "uint64": _T.TYPE_UINT64,
           "bool": _T.TYPE_BOOL, "float": _T.TYPE_FLOAT, "double": _T.TYPE_DOUBLE, "string": _T.TYPE_STRING}


_TOP_LEVEL_ENUMS = ("Phase",)


def _build():
    fd = descriptor_pb2.FileDescriptorProto(name="shf_caffe.proto", package="caffe", syntax="proto2")
    # `Phase` is a file-level enum in caffe.proto; every other enum is nested in the message that uses it (Engine is
    # declared once per layer parameter, PoolMethod.MAX and EltwiseOp.MAX would collide at file level)
    for ename in _TOP_LEVEL_ENUMS:
        e = fd.enum_type.add(name=ename)
        for k, v in ENUMS[ename].items():
            e.value.add(name=k, number=v)
    for mname, fields in SCHEMA.items():
        m = fd.message_type.add(name=mname)
        for ename in sorted({typ[5:] for _, typ, _ in fields.values() if typ.startswith("enum:")} - set(_TOP_LEVEL_ENUMS)):
            e = m.enum_type.add(name=ename)
            for k, v in ENUMS[ename].items():
                e.value.add(name=k, number=v)
        for fname, (num, typ, lab) in sorted(fields.items(), key=lambda kv: kv[1][0]):
            f = m.field.add(name=fname, number=num,
                            label=_T.LABEL_OPTIONAL if lab == "o" else _T.LABEL_REPEATED)
            if typ in SCHEMA:
                f.type, f.type_name = _T.TYPE_MESSAGE, ".caffe." + typ
            elif typ.startswith("enum:"):
                nested = typ[5:] not in _TOP_LEVEL_ENUMS
                f.type, f.type_name = _T.TYPE_ENUM, ".caffe." + (mname + "." if nested else "") + typ[5:]
            else:
                f.type = _SCALAR[typ]
            if lab == "p":
                f.options.packed = True
            d = DEFAULTS.get((mname, fname))
            if d is not None and lab == "o" and typ not in SCHEMA:
                if typ.startswith("enum:"):
                    f.default_value = [k for k, v in ENUMS[typ[5:]].items() if v == d][0]
                elif typ == "bool":
                    f.default_value = "true" if d else "false"
                else:
                    f.default_value = str(d)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {m: message_factory.GetMessageClass(pool.FindMessageTypeByName("caffe." + m)) for m in SCHEMA}, pool


_CLASSES, _POOL = _build()
globals().update(_CLASSES)
TRAIN, TEST = 0, 1
