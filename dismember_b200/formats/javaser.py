"""Reader for the Java Object Serialization stream (magic 0xACED0005).

The reference saves models with ``java.io.ObjectOutputStream``
(tdm/src/main/scala/com/mass/tdm/utils/Serialization.scala:60-101,
deep-retrieval/src/main/scala/com/mass/dr/model/DeepRetrieval.scala:71-106).
No JVM exists in the build image, so the weights are pulled out of those
files by walking the stream grammar directly (JDK "Object Serialization
Stream Protocol", chapter 6).  Only what the model files need is supported:
class descriptors, objects (default + writeObject annotation), primitive and
object arrays, strings, enums, references, block data.

Primitive arrays are returned as numpy arrays (big-endian converted to native);
objects as ``JavaObject`` with ``.classname`` and ``.fields`` (dict) and
``.annotations`` (list of raw blockdata/objects written by custom writeObject).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

TC_NULL, TC_REFERENCE, TC_CLASSDESC, TC_OBJECT, TC_STRING, TC_ARRAY = 0x70, 0x71, 0x72, 0x73, 0x74, 0x75
TC_CLASS, TC_BLOCKDATA, TC_ENDBLOCKDATA, TC_RESET, TC_BLOCKDATALONG = 0x76, 0x77, 0x78, 0x79, 0x7A
TC_EXCEPTION, TC_LONGSTRING, TC_PROXYCLASSDESC, TC_ENUM = 0x7B, 0x7C, 0x7D, 0x7E
BASE_HANDLE = 0x7E0000
SC_WRITE_METHOD, SC_SERIALIZABLE, SC_EXTERNALIZABLE, SC_BLOCK_DATA = 0x01, 0x02, 0x04, 0x08

_PRIM = {
    "B": (">i1", 1), "C": (">u2", 2), "D": (">f8", 8), "F": (">f4", 4),
    "I": (">i4", 4), "J": (">i8", 8), "S": (">i2", 2), "Z": ("?", 1),
}


@dataclass
class ClassDesc:
    name: str
    flags: int
    fields: List[tuple]            # (typecode, name, classname-or-None)
    superclass: Optional["ClassDesc"]

    def hierarchy(self) -> List["ClassDesc"]:
        out, c = [], self
        while c is not None:
            out.append(c)
            c = c.superclass
        return out[::-1]           # super-most first, as written on the wire


@dataclass
class JavaObject:
    classname: str
    fields: Dict[str, Any] = field(default_factory=dict)
    annotations: List[Any] = field(default_factory=list)

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return f"<JavaObject {self.classname} fields={list(self.fields)}>"


@dataclass
class JavaEnum:
    classname: str
    constant: str


class JavaSerError(ValueError):
    pass


class _Reader:
    def __init__(self, data: bytes):
        self.b = data
        self.p = 0
        self.handles: List[Any] = []
        self.spans: List[tuple] = []   # (byte offset of the payload, element typecode, element count) per primitive array, stream order

    def u1(self):
        v = self.b[self.p]; self.p += 1; return v

    def take(self, n):
        v = self.b[self.p:self.p + n]
        if len(v) != n:
            raise JavaSerError("truncated stream")
        self.p += n
        return v

    def unpack(self, fmt):
        n = struct.calcsize(fmt)
        return struct.unpack(fmt, self.take(n))[0]

    def utf(self):
        n = self.unpack(">H")
        return self.take(n).decode("utf-8", errors="replace")

    def new_handle(self, obj):
        self.handles.append(obj)
        return len(self.handles) - 1

    # ---- grammar -------------------------------------------------------
    def content(self):
        """object | blockdata ; returns python value"""
        tc = self.b[self.p]
        if tc == TC_BLOCKDATA:
            self.p += 1
            n = self.u1()
            return bytes(self.take(n))
        if tc == TC_BLOCKDATALONG:
            self.p += 1
            n = self.unpack(">i")
            return bytes(self.take(n))
        return self.object()

    def object(self):
        tc = self.u1()
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            h = self.unpack(">i") - BASE_HANDLE
            if not 0 <= h < len(self.handles):
                raise JavaSerError(f"bad handle {h} at {self.p}")
            return self.handles[h]
        if tc == TC_CLASSDESC or tc == TC_PROXYCLASSDESC:
            self.p -= 1
            return self.class_desc()
        if tc == TC_CLASS:
            cd = self.class_desc()
            self.new_handle(cd)
            return cd
        if tc == TC_STRING:
            h = self.new_handle(None)
            s = self.utf()
            self.handles[h] = s
            return s
        if tc == TC_LONGSTRING:
            h = self.new_handle(None)
            n = self.unpack(">q")
            s = self.take(n).decode("utf-8", errors="replace")
            self.handles[h] = s
            return s
        if tc == TC_ARRAY:
            return self.array()
        if tc == TC_OBJECT:
            return self.new_object()
        if tc == TC_ENUM:
            cd = self.class_desc()
            e = JavaEnum(cd.name, "")
            self.new_handle(e)
            e.constant = self.object()
            return e
        if tc == TC_RESET:
            self.handles.clear()
            return None
        raise JavaSerError(f"unsupported type code 0x{tc:02x} at {self.p - 1}")

    def class_desc(self) -> Optional[ClassDesc]:
        tc = self.u1()
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            h = self.unpack(">i") - BASE_HANDLE
            cd = self.handles[h]
            if not isinstance(cd, ClassDesc):
                raise JavaSerError("reference is not a class descriptor")
            return cd
        if tc == TC_PROXYCLASSDESC:
            cd = ClassDesc("<proxy>", SC_SERIALIZABLE, [], None)
            self.new_handle(cd)
            n = self.unpack(">i")
            for _ in range(n):
                self.utf()
            self.annotation()
            cd.superclass = self.class_desc()
            return cd
        if tc != TC_CLASSDESC:
            raise JavaSerError(f"expected class desc, got 0x{tc:02x} at {self.p - 1}")
        name = self.utf()
        self.take(8)  # serialVersionUID
        cd = ClassDesc(name, 0, [], None)
        self.new_handle(cd)
        cd.flags = self.u1()
        nf = self.unpack(">H")
        for _ in range(nf):
            t = chr(self.u1())
            fname = self.utf()
            cname = None
            if t in "[L":
                cname = self.object()
            cd.fields.append((t, fname, cname))
        self.annotation()
        cd.superclass = self.class_desc()
        return cd

    def annotation(self) -> list:
        out = []
        while self.b[self.p] != TC_ENDBLOCKDATA:
            out.append(self.content())
        self.p += 1
        return out

    def array(self):
        cd = self.class_desc()
        h = self.new_handle(None)
        n = self.unpack(">i")
        et = cd.name[1]
        if et in _PRIM:
            dt, sz = _PRIM[et]
            self.spans.append((self.p, et, n))
            raw = self.take(n * sz)
            arr = np.frombuffer(raw, dtype=dt).astype(np.dtype(dt).newbyteorder("="))
            self.handles[h] = arr
            return arr
        lst: list = []
        self.handles[h] = lst
        for _ in range(n):
            lst.append(self.object())
        return lst

    def new_object(self):
        cd = self.class_desc()
        obj = JavaObject(cd.name)
        self.new_handle(obj)
        for c in cd.hierarchy():
            if c.flags & SC_SERIALIZABLE:
                for t, fname, _ in c.fields:
                    if t in _PRIM:
                        fmt = {"B": ">b", "C": ">H", "D": ">d", "F": ">f", "I": ">i",
                               "J": ">q", "S": ">h", "Z": ">?"}[t]
                        obj.fields[fname] = self.unpack(fmt)
                    else:
                        obj.fields[fname] = self.object()
                if c.flags & SC_WRITE_METHOD:
                    obj.annotations.extend(self.annotation())
            elif c.flags & SC_EXTERNALIZABLE:
                if c.flags & SC_BLOCK_DATA:
                    obj.annotations.extend(self.annotation())
                else:
                    raise JavaSerError("externalizable without block data is not parseable")
        return obj


def load(data: bytes) -> List[Any]:
    """Parse a whole stream; returns the list of top-level contents."""
    r = _Reader(data)
    if r.unpack(">H") != 0xACED or r.unpack(">H") != 5:
        raise JavaSerError("not a Java serialization stream")
    out = []
    while r.p < len(r.b):
        out.append(r.content())
    return out


def load_file(path: str) -> List[Any]:
    with open(path, "rb") as f:
        return load(f.read())


def primitive_array_spans(data: bytes) -> List[tuple]:
    """(byte offset of the payload, element typecode, element count) of every primitive array, in stream order."""
    r = _Reader(data)
    if r.unpack(">H") != 0xACED or r.unpack(">H") != 5:
        raise JavaSerError("not a Java serialization stream")
    while r.p < len(r.b):
        r.content()
    return list(r.spans)


def inject_weights(data: bytes, params, min_len: int = 1000, which: int = 0) -> bytes:
    """The injector of SURVEY 8f rank 1: a copy of a Java-serialised model whose parameter storage holds `params`.

    `Module.parameters()` of a saved model is ONE primitive array shared by every tensor (compact storage made by
    `adjustParameters`, scalann/.../nn/abstractnn/AbstractModule.scala:163) -- the first `[F` / `[D` array of at least
    `min_len` elements in the stream (tools/make_golden.py reads the weights from the same place; the array after it is
    the gradient buffer).  Only its payload bytes change, so `Serialization.loadModel`
    (tdm/.../utils/Serialization.scala:81-101) reads the file as before and finds the new weights.  `which` picks a
    later qualifying array instead (a Deep Retrieval file holds several storages: see `primitive_array_spans`)."""
    params = np.asarray(params)
    seen = 0
    for off, et, n in primitive_array_spans(data):
        if et not in ("F", "D") or n < min_len:
            continue
        seen += 1
        if seen <= which:
            continue
        if n != params.size:
            raise JavaSerError(f"the model's parameter array holds {n} values, got {params.size}")
        dt, sz = _PRIM[et]
        if (et == "F") != (params.dtype == np.float32):
            raise JavaSerError(f"the model stores {'Float' if et == 'F' else 'Double'} parameters, got {params.dtype}")
        raw = np.ascontiguousarray(params.ravel()).astype(dt).tobytes()
        return data[:off] + raw + data[off + n * sz:]
    raise JavaSerError("no parameter array found")


def inject_weights_file(src: str, dst: str, params, min_len: int = 1000, which: int = 0) -> None:
    with open(src, "rb") as f:
        data = f.read()
    with open(dst, "wb") as f:
        f.write(inject_weights(data, params, min_len, which))


def walk(obj, fn, _seen=None, _path="$"):
    """Depth-first visit of every JavaObject / array reachable from obj."""
    if _seen is None:
        _seen = set()
    if id(obj) in _seen:
        return
    if isinstance(obj, JavaObject):
        _seen.add(id(obj))
        fn(_path, obj)
        for k, v in obj.fields.items():
            walk(v, fn, _seen, f"{_path}.{k}")
        for i, v in enumerate(obj.annotations):
            walk(v, fn, _seen, f"{_path}@{i}")
    elif isinstance(obj, np.ndarray):
        _seen.add(id(obj))
        fn(_path, obj)
    elif isinstance(obj, list):
        _seen.add(id(obj))
        for i, v in enumerate(obj):
            walk(v, fn, _seen, f"{_path}[{i}]")


def primitive_arrays(objs, min_len: int = 1):
    """All distinct primitive arrays in stream order of first reach: [(path, ndarray)]."""
    found = []

    def fn(path, o):
        if isinstance(o, np.ndarray) and o.size >= min_len:
            found.append((path, o))

    for i, o in enumerate(objs):
        walk(o, fn, None, f"$[{i}]")
    return found
