"""Host-side tree / alignment readers and flatteners (the data formats either side of the hot path).

Mirrors what the reference's Java host obtains from PAL before it schedules columns
(`/root/reference/src/subrecon/SubRecon.java:326-355`): a rooted tree whose children keep Newick
order, and an alignment addressed by sequence name. Behaviour follows SURVEY.md Appendix B.6/B.7
(PAL `ReadTree::readNH`, `FormattedInput::readLabel`, `AlignmentReaders::readFastaSequences`,
`SimpleAlignment`, `AminoAcids::getStateImpl`). The output is the flattened form the C ABI takes
(`include/subrecon_b200.h`: CSR tree, `uint8 states[taxon][site]`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

AA_ORDER = "ARNDCQEGHILKMFPSTWYV"          # PAL state order, pal/datatype/AminoAcids.class::<clinit>
MISSING = 255                               # anything outside 0..19 is missing data (JointBranchReconstruction.java:199-205)

_STATE_LUT = np.full(256, MISSING, dtype=np.uint8)
for _i, _c in enumerate(AA_ORDER):
    _STATE_LUT[ord(_c)] = _i
    _STATE_LUT[ord(_c.lower())] = _i


@dataclass
class Tree:
    """Rooted tree; node 0 is the root; `children[v]` keeps Newick order; `blen[v]` = edge to the parent."""
    children: list = field(default_factory=list)
    blen: list = field(default_factory=list)
    name: list = field(default_factory=list)

    def add(self, name="", blen=0.0):
        self.children.append([])
        self.blen.append(float(blen))
        self.name.append(name)
        return len(self.blen) - 1

    @property
    def n_nodes(self):
        return len(self.blen)

    def is_leaf(self, v):
        return not self.children[v]

    def leaves(self):
        return [v for v in range(self.n_nodes) if not self.children[v]]

    def to_newick(self):
        out = {}
        order = []
        stack = [0]
        while stack:
            v = stack.pop()
            order.append(v)
            stack.extend(self.children[v])
        for v in reversed(order):
            s = self.name[v] if not self.children[v] else "(" + ",".join(out.pop(c) for c in self.children[v]) + ")"
            out[v] = s if v == 0 else f"{s}:{self.blen[v]!r}"
        return out[0] + ";"


_LABEL_STOP = set(" \t\r\n:,;()[]{}")


def parse_newick(text: str) -> Tree:
    """Iterative Newick reader with PAL's conventions: missing length = 0.0 (clamped to 1e-9 later by
    PAL's `setDistance`), a node with exactly one child is an error, internal labels are ignored."""
    t = Tree()
    pos = 0
    n = len(text)

    def skip_ws():
        nonlocal pos
        while pos < n and text[pos] in " \t\r\n":
            pos += 1

    root = t.add()
    cur = root
    stack = []
    skip_ws()
    if pos >= n or text[pos] != "(":
        raise ValueError("Missing tree")
    # state machine: at each step we are about to read a subtree for a new child of `cur`, or closing
    expect_child = False
    pos += 1
    stack.append(cur)
    expect_child = True
    while stack:
        skip_ws()
        if pos >= n:
            raise ValueError("Unexpected end of tree")
        c = text[pos]
        if expect_child:
            v = t.add()
            t.children[stack[-1]].append(v)
            if c == "(":
                pos += 1
                stack.append(v)
                continue
            # leaf label
            s = pos
            while pos < n and text[pos] not in _LABEL_STOP:
                pos += 1
            t.name[v] = text[s:pos]
            done = v
        else:
            # just closed node `done`; an optional internal label follows
            s = pos
            while pos < n and text[pos] not in _LABEL_STOP:
                pos += 1
            done = closed
        skip_ws()
        if pos < n and text[pos] == ":":
            pos += 1
            skip_ws()
            s = pos
            while pos < n and text[pos] not in _LABEL_STOP:
                pos += 1
            t.blen[done] = float(text[s:pos])
            skip_ws()
        if done == root:
            break
        if pos >= n:
            raise ValueError("Unexpected end of tree")
        c = text[pos]
        if c == ",":
            pos += 1
            expect_child = True
        elif c == ")":
            pos += 1
            closed = stack.pop()
            if len(t.children[closed]) == 1:
                raise ValueError("Node with single child enountered")
            expect_child = False
        else:
            raise ValueError(f"Missing ',' or ')' at {pos}")
    return t


def read_tree(path) -> Tree:
    with open(path) as fh:
        return parse_newick(fh.read())


def read_fasta(path):
    """-> (names, sequences). Names are the full header after '>' trimmed; sequence lines lose ' \\r\\n\\t'."""
    names, seqs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                names.append(line[1:].strip())
                seqs.append([])
            elif names:
                seqs[-1].append(line.translate({32: None, 9: None, 10: None, 13: None}))
    return names, ["".join(s) for s in seqs]


def read_phylip(path):
    """Sequential / interleaved PHYLIP with names separated from data by whitespace."""
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh if ln.strip()]
    n, length = (int(x) for x in lines[0].split()[:2])
    names, seqs = [], []
    for i in range(n):
        parts = lines[1 + i].split(None, 1)
        names.append(parts[0])
        seqs.append(parts[1].replace(" ", "") if len(parts) > 1 else "")
    k = 0
    for ln in lines[1 + n:]:
        seqs[k % n] += ln.replace(" ", "")
        k += 1
    return names, [s[:length] if len(s) >= length else s for s in seqs]


def encode_alignment(seqs) -> np.ndarray:
    """chars -> uint8 states[taxon][site]; sequences are right-padded with '-' to the longest
    (`SimpleAlignment::getPadded`); letters map case-insensitively to 0..19, everything else to MISSING."""
    length = max((len(s) for s in seqs), default=0)
    out = np.full((len(seqs), length), MISSING, dtype=np.uint8)
    for i, s in enumerate(seqs):
        if s:
            out[i, :len(s)] = _STATE_LUT[np.frombuffer(s.encode("latin-1", "replace"), dtype=np.uint8)]
    return out


def alignment_chars(seqs) -> np.ndarray:
    """chars -> uint8 letters[taxon][site] for the library's "input_chars" mode: the same right-padding with '-' as
    `encode_alignment`, but the translation to states is left to the device (one pass over the uploaded bytes)."""
    length = max((len(s) for s in seqs), default=0)
    out = np.full((len(seqs), length), ord("-"), dtype=np.uint8)
    for i, s in enumerate(seqs):
        if s:
            out[i, :len(s)] = np.frombuffer(s.encode("latin-1", "replace"), dtype=np.uint8)
    return out


@dataclass
class FlatTree:
    n_nodes: int
    child_off: np.ndarray     # int32 [n_nodes+1]
    child_idx: np.ndarray     # int32 [n_edges]
    blen: np.ndarray          # float64 [n_nodes]
    leaf_row: np.ndarray      # int32 [n_nodes], -1 for internal nodes
    root: int


def flatten_tree(tree: Tree, row_of_name: dict | None = None) -> FlatTree:
    """CSR form for the C ABI. `row_of_name` maps tip names to alignment rows (default: tips in node order).
    An unknown tip name raises — in the reference it makes *every* column throw."""
    nn = tree.n_nodes
    off = np.zeros(nn + 1, dtype=np.int32)
    for v in range(nn):
        off[v + 1] = off[v] + len(tree.children[v])
    idx = np.fromiter((c for v in range(nn) for c in tree.children[v]), dtype=np.int32, count=int(off[-1]))
    leaf_row = np.full(nn, -1, dtype=np.int32)
    r = 0
    for v in range(nn):
        if not tree.children[v]:
            if row_of_name is None:
                leaf_row[v] = r
                r += 1
            else:
                if tree.name[v] not in row_of_name:
                    raise KeyError(f"tree tip {tree.name[v]!r} is not in the alignment")
                leaf_row[v] = row_of_name[tree.name[v]]
    return FlatTree(nn, off, idx, np.asarray(tree.blen, dtype=np.float64), leaf_row, 0)
