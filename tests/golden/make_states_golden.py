"""Generate tests/golden/aa_states_golden.json by EXECUTING the reference's tip-state lookup from PAL 1.51 bytecode:

    DataTypeTool.getUniverisalAminoAcids().getState(c)        (molevo/AdvancedAlignmentAminoAcid.java:34-40)

for every byte value c = 0..255 (PAL's `SimpleAlignment.getData` returns the residue letter as a char). The reference
treats a result outside 0..19 as missing data (recon/JointBranchReconstruction.java:199-205). Build container only
(needs /root/reference/lib/pal.jar); the jar runs under oracle/jvm/interp.py.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jvm.interp import VM  # noqa: E402

JAR = "/root/reference/lib/pal.jar"


def jar_states(jar=JAR):
    vm = VM(jar)
    dt = vm.call("pal/datatype/DataTypeTool", "getUniverisalAminoAcids", "()Lpal/datatype/MolecularDataType;")
    return dt.cls, [int(vm.call(dt, "getState", "(C)I", c)) for c in range(256)], sorted(vm.stubbed)


def main():
    cls, states, stubbed = jar_states(sys.argv[1] if len(sys.argv) > 1 else JAR)
    out = {"_provenance": "pal.jar (PAL 1.51) bytecode under oracle/jvm/interp.py: DataTypeTool.getUniverisalAminoAcids() -> "
                          + cls + ".getState(char) for char 0..255", "states": states, "_stubbed_host_calls": stubbed}
    with open(os.path.join(ROOT, "tests/golden/aa_states_golden.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote aa_states_golden.json:", sum(0 <= s < 20 for s in states), "bytes map to a state")


if __name__ == "__main__":
    main()
