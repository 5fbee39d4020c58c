"""Extract the PAL 1.51 amino-acid model tables from the reference's own jar.

TEST INFRASTRUCTURE (run in the build container only; /root/reference is absent on the GPU box).
Reads `lib/pal.jar!pal/substmodel/{WAG,JTT,Dayhoff,BLOSUM62}.class`:
  * `rebuildRateMatrix([[D[D)V`  — 190 stores `rate[i][j] = const` (strict upper triangle)
  * `getOriginalFrequencies([D)V` — 20 stores `f[i] = const`
by pattern-matching the straight-line bytecode (aload_1; push i; aaload; push j; ldc2_w; dastore).
Writes `oracle/pal_models.json` (committed) — the same tables SURVEY.md Appendix C lists.

Usage: python oracle/jvm/extract_pal_tables.py [/root/reference/lib/pal.jar]
"""
import json
import os
import struct
import sys

try:
    from .classfile import Jar
except ImportError:                                   # run as a script
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle.jvm.classfile import Jar  # noqa: E402


def _scan(cf, code):
    """Yield ('int', v) | ('dbl', v) | ('op', name) for the small opcode subset these methods use."""
    pc = 0
    while pc < len(code):
        op = code[pc]
        if 2 <= op <= 8:
            yield ("int", op - 3); pc += 1
        elif op == 0x10:
            yield ("int", struct.unpack_from(">b", code, pc + 1)[0]); pc += 2
        elif op == 0x11:
            yield ("int", struct.unpack_from(">h", code, pc + 1)[0]); pc += 3
        elif op == 0x14:
            ix = struct.unpack_from(">H", code, pc + 1)[0]
            assert cf.cp[ix][0] == "Double"
            yield ("dbl", cf.cp[ix][1]); pc += 3
        elif op in (0x0E, 0x0F):
            yield ("dbl", float(op - 0x0E)); pc += 1
        elif op == 0x2A:
            yield ("op", "aload_0"); pc += 1
        elif op == 0x2B:
            yield ("op", "aload_1"); pc += 1
        elif op == 0x32:
            yield ("op", "aaload"); pc += 1
        elif op == 0x52:
            yield ("op", "dastore"); pc += 1
        elif op == 0xB1:
            yield ("op", "return"); pc += 1
        else:
            raise ValueError(f"unexpected opcode 0x{op:02x} at pc {pc}")


def extract(jar, cls):
    cf = jar.load(cls)
    q = {}
    toks = list(_scan(cf, cf.methods[("rebuildRateMatrix", "([[D[D)V")].code))
    i = 0
    while i < len(toks):
        if toks[i] == ("op", "return"):
            break
        assert toks[i] == ("op", "aload_1"), toks[i]
        (_, r), aal, (_, c), (kd, v), st = toks[i + 1:i + 6]
        assert aal == ("op", "aaload") and st == ("op", "dastore") and kd == "dbl"
        q[(r, c)] = v
        i += 6
    assert len(q) == 190 and all(r < c for r, c in q), len(q)
    rows = [[q[(r, c)] for c in range(r + 1, 20)] for r in range(19)]
    f = {}
    toks = list(_scan(cf, cf.methods[("getOriginalFrequencies", "([D)V")].code))
    i = 0
    while i < len(toks):
        if toks[i] == ("op", "return"):
            break
        assert toks[i] == ("op", "aload_0"), toks[i]
        (_, ix), (kd, v), st = toks[i + 1:i + 4]
        assert st == ("op", "dastore") and kd == "dbl"
        f[ix] = v
        i += 4
    assert sorted(f) == list(range(20))
    return {"freq": [f[k] for k in range(20)], "q_upper_rows": rows}


def main():
    jar_path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/lib/pal.jar"
    jar = Jar(jar_path)
    out = {"_provenance": "constants of lib/pal.jar (PAL 1.51) pal/substmodel/*.class::rebuildRateMatrix / "
                          "getOriginalFrequencies, extracted by oracle/jvm/extract_pal_tables.py; state order "
                          "A R N D C Q E G H I L K M F P S T W Y V"}
    for name in ("WAG", "JTT", "Dayhoff", "BLOSUM62"):
        out[name] = extract(jar, "pal/substmodel/" + name)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pal_models.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    for name in ("WAG", "JTT", "Dayhoff", "BLOSUM62"):
        print(name, "sum q = %.10f" % sum(sum(r) for r in out[name]["q_upper_rows"]), "sum f = %.17g" % sum(out[name]["freq"]))


if __name__ == "__main__":
    main()
