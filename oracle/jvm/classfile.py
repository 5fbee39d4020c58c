"""Minimal JVM class-file reader (constant pool, fields, methods, Code attributes).

TEST INFRASTRUCTURE. Used by `oracle/jvm/interp.py` (a small bytecode interpreter) and
`oracle/jvm/extract_pal_tables.py` to read the *reference's own* third-party numerics out of
`/root/reference/lib/pal.jar` (PAL 1.51, bytecode only, no source). Nothing here is shipped in
the product path. Format follows the JVM specification, chapter 4 (class version 46).
"""
from __future__ import annotations

import struct
import zipfile
from dataclasses import dataclass, field


@dataclass
class Method:
    name: str
    desc: str
    access: int
    max_stack: int = 0
    max_locals: int = 0
    code: bytes = b""
    exc_table: list = field(default_factory=list)


@dataclass
class ClassFile:
    name: str
    super_name: str | None
    interfaces: list
    cp: list
    fields: list          # (access, name, desc, constant_value_index|None)
    methods: dict         # (name, desc) -> Method
    access: int = 0


def parse_class(data: bytes) -> ClassFile:
    pos = 0

    def u1():
        nonlocal pos
        v = data[pos]; pos += 1; return v

    def u2():
        nonlocal pos
        v = struct.unpack_from(">H", data, pos)[0]; pos += 2; return v

    def u4():
        nonlocal pos
        v = struct.unpack_from(">I", data, pos)[0]; pos += 4; return v

    assert u4() == 0xCAFEBABE
    u2(); u2()
    n = u2()
    cp = [None] * n
    i = 1
    while i < n:
        tag = u1()
        if tag == 1:
            ln = u2(); cp[i] = ("Utf8", data[pos:pos + ln].decode("utf-8", "replace")); pos += ln
        elif tag == 3:
            cp[i] = ("Integer", struct.unpack_from(">i", data, pos)[0]); pos += 4
        elif tag == 4:
            cp[i] = ("Float", struct.unpack_from(">f", data, pos)[0]); pos += 4
        elif tag == 5:
            cp[i] = ("Long", struct.unpack_from(">q", data, pos)[0]); pos += 8; i += 1
        elif tag == 6:
            cp[i] = ("Double", struct.unpack_from(">d", data, pos)[0]); pos += 8; i += 1
        elif tag == 7:
            cp[i] = ("Class", u2())
        elif tag == 8:
            cp[i] = ("String", u2())
        elif tag in (9, 10, 11):
            cp[i] = ({9: "Fieldref", 10: "Methodref", 11: "InterfaceMethodref"}[tag], u2(), u2())
        elif tag == 12:
            cp[i] = ("NameAndType", u2(), u2())
        else:
            raise ValueError(f"constant-pool tag {tag}")
        i += 1

    def utf(ix):
        return cp[ix][1]

    access = u2()
    this_name = utf(cp[u2()][1])
    sup = u2()
    super_name = utf(cp[sup][1]) if sup else None
    interfaces = [utf(cp[u2()][1]) for _ in range(u2())]

    def attributes():
        out = []
        for _ in range(u2()):
            nm = utf(u2()); ln = u4()
            nonlocal pos
            out.append((nm, data[pos:pos + ln])); pos += ln
        return out

    fields = []
    for _ in range(u2()):
        acc = u2(); nm = utf(u2()); ds = utf(u2())
        cv = None
        for an, ab in attributes():
            if an == "ConstantValue":
                cv = struct.unpack(">H", ab)[0]
        fields.append((acc, nm, ds, cv))

    methods = {}
    for _ in range(u2()):
        acc = u2(); nm = utf(u2()); ds = utf(u2())
        m = Method(nm, ds, acc)
        for an, ab in attributes():
            if an == "Code":
                ms, ml, cl = struct.unpack_from(">HHI", ab, 0)
                m.max_stack, m.max_locals = ms, ml
                m.code = ab[8:8 + cl]
                p = 8 + cl
                ne = struct.unpack_from(">H", ab, p)[0]; p += 2
                for _e in range(ne):
                    m.exc_table.append(struct.unpack_from(">HHHH", ab, p)); p += 8
        methods[(nm, ds)] = m
    return ClassFile(this_name, super_name, interfaces, cp, fields, methods, access)


class Jar:
    """Lazy class loader over a jar."""

    def __init__(self, path):
        self.zip = zipfile.ZipFile(path)
        self.cache = {}

    def has(self, name):
        try:
            self.zip.getinfo(name + ".class"); return True
        except KeyError:
            return False

    def load(self, name) -> ClassFile:
        if name not in self.cache:
            self.cache[name] = parse_class(self.zip.read(name + ".class"))
        return self.cache[name]


# ---- helpers to resolve constant-pool references
def cp_class(cf, ix):
    return cf.cp[cf.cp[ix][1]][1]


def cp_member(cf, ix):
    _, ci, nti = cf.cp[ix]
    cls = cp_class(cf, ci)
    _, ni, di = cf.cp[nti]
    return cls, cf.cp[ni][1], cf.cp[di][1]


def cp_string(cf, ix):
    return cf.cp[cf.cp[ix][1]][1]
