"""DBoW2 vocabulary as flat arrays (the layout orbv_create takes) and the device transform (host-side mirror of ORBVocabulary / Frame::ComputeBoW).

Flat layout: node 0 = root; the children of node i are child_ids[child_start[i]:child_start[i+1]] in Node::children order; a node without
children is a word (word_id, weight).  load_text reads the reference's ORBvoc.txt format (TemplatedVocabulary.h:1338-1419): a header line
`k L scoring weighting`, then one line per node `parent is_leaf d0 .. d31 weight`, node ids counting from 1 in file order, word ids
counting the leaves in file order."""
import ctypes

import numpy as np

from . import load, _check, _ptr


def from_nodes(k, L, parent, is_leaf, node_desc, weight):
    """parent[i], is_leaf[i], node_desc[i], weight[i] for nodes 1..n-1 in file order (index 0 = root placeholder)."""
    parent = np.asarray(parent, np.int64); n = len(parent)
    order = np.argsort(parent[1:], kind="stable") + 1                 # children grouped by parent, file order inside a group
    counts = np.bincount(parent[1:], minlength=n)
    child_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    leaf = np.asarray(is_leaf, bool).copy(); leaf[0] = False
    word_id = np.full(n, -1, np.int32); word_id[leaf] = np.arange(int(leaf.sum()), dtype=np.int32)
    return dict(k=int(k), L=int(L), node_desc=np.ascontiguousarray(node_desc, np.uint8), child_start=child_start, child_ids=order.astype(np.int32),
                word_id=word_id, weight=np.ascontiguousarray(weight, np.float64), parent=parent.astype(np.int32), is_leaf=leaf)


def save_text(vocab, path):
    """Write the ORBvoc.txt format (TemplatedVocabulary::saveToTextFile, TemplatedVocabulary.h:1422-1447).  No trailing newline: the reference's
    loader reads one more (empty) node line after a final newline (`while(!f.eof()) getline`, :1377-1380)."""
    n = len(vocab["word_id"])
    lines = ["%d %d 0 0" % (vocab["k"], vocab["L"])]
    for i in range(1, n):
        lines.append("%d %d %s %.17g" % (vocab["parent"][i], 1 if vocab["is_leaf"][i] else 0, " ".join(str(int(b)) for b in vocab["node_desc"][i]), vocab["weight"][i]))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def load_text(path):
    """Reads the nodes the file lists.  Divergence from the reference's loader, kept on purpose: after a final newline `while(!f.eof()) getline`
    (TemplatedVocabulary.h:1377-1410) appends one more node under the root whose leaf flag, weight and descriptor are whatever the failed stream
    extractions leave behind (uninitialised memory in a real OpenCV build).  It cannot be reproduced; modelling it as a zero-descriptor, weight-0 node
    was tried and DROPS 45 of 1956 words of a KITTI-shape frame on the shipped ORBvoc.txt, whereas the reference's own object code (oracle/_ref) gives
    the same BowVector with and without the final newline -- which is what this loader gives (tests/test_oracle_vs_reference.py)."""
    with open(path) as f:
        k, L, n1, n2 = [int(x) for x in f.readline().split()[:4]]
        if n1 != 0 or n2 != 0:
            raise ValueError("only TF_IDF weighting with L1 scoring (the configuration ORBSLAMM ships) is supported")
        rows = np.loadtxt(f, dtype=np.float64)
    n = len(rows) + 1
    parent = np.zeros(n, np.int64); parent[1:] = rows[:, 0].astype(np.int64)
    is_leaf = np.zeros(n, bool); is_leaf[1:] = rows[:, 1] > 0
    desc = np.zeros((n, 32), np.uint8); desc[1:] = rows[:, 2:34].astype(np.uint8)
    weight = np.zeros(n, np.float64); weight[1:] = rows[:, 34]
    return from_nodes(k, L, parent, is_leaf, desc, weight)


def synthetic(k=10, L=3, seed=0, stop_fraction=0.02, tie_fraction=0.05):
    """A random complete k-ary tree of depth L in file (depth-first) order, like DBoW2 writes its vocabularies; a few words are stopped
    (weight 0) and a few siblings share a descriptor so that the first-minimum rule of the tree walk matters."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [0], [False], [np.zeros(32, np.uint8)], [0.0]

    def grow(pid, level, centre):
        prev = None
        for _ in range(k):
            d = centre ^ np.packbits(rng.random(256) < (0.5 / (level + 1))).astype(np.uint8)
            if prev is not None and rng.random() < tie_fraction:
                d = prev.copy()
            prev = d
            nid = len(parent)
            parent.append(pid); leaf.append(level == L); desc.append(d)
            weight.append(0.0 if (level == L and rng.random() < stop_fraction) else (float(rng.uniform(0.5, 9.0)) if level == L else 0.0))
            if level < L:
                grow(nid, level + 1, d)

    grow(0, 1, np.zeros(32, np.uint8))
    return from_nodes(k, L, parent, leaf, np.stack(desc), weight)


class ORBVocabulary:
    """Device-resident vocabulary; transform() = ORBVocabulary::transform(vDesc, BowVec, FeatVec, levelsup) for a batch of frames."""

    def __init__(self, vocab, device=0):
        self._L = load()
        L = self._L
        vp, i = ctypes.c_void_p, ctypes.c_int
        L.orbv_create.argtypes = [ctypes.POINTER(vp), i, i, i, i, vp, vp, vp, vp, vp]; L.orbv_create.restype = i
        L.orbv_destroy.argtypes = [vp]; L.orbv_destroy.restype = i
        L.orbv_transform.argtypes = [vp, i, vp, vp, i, i] + [vp] * 9 + [i]; L.orbv_transform.restype = i
        self.vocab = vocab
        self._h = vp()
        _check(L.orbv_create(ctypes.byref(self._h), int(device), vocab["k"], vocab["L"], len(vocab["word_id"]), _ptr(vocab["node_desc"]), _ptr(vocab["child_start"]),
                             _ptr(vocab["child_ids"]), _ptr(vocab["word_id"]), _ptr(vocab["weight"])))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.orbv_destroy(h)

    def transform(self, desc, counts, levelsup=4):
        """desc u8[n_frames, slab, 32].  Returns a list (one per frame) of dict(word_of, node_of, bow_ids, bow_vals, fv=dict(nodes, start, items))."""
        d = np.ascontiguousarray(desc, np.uint8); n, slab = d.shape[0], d.shape[1]
        cnt = np.ascontiguousarray(counts, np.int32)
        wo = np.zeros((n, slab), np.int32); no = np.zeros((n, slab), np.int32); bi = np.zeros((n, slab), np.int32); bv = np.zeros((n, slab), np.float64)
        bc = np.zeros(n, np.int32); fn = np.zeros((n, slab), np.int32); fs = np.zeros((n, slab + 1), np.int32); fi = np.zeros((n, slab), np.int32); fc = np.zeros(n, np.int32)
        _check(self._L.orbv_transform(self._h, n, _ptr(d), _ptr(cnt), slab, int(levelsup), _ptr(wo), _ptr(no), _ptr(bi), _ptr(bv), _ptr(bc), _ptr(fn), _ptr(fs), _ptr(fi),
                                      _ptr(fc), 0))
        out = []
        for f in range(n):
            nf = fc[f]
            out.append(dict(word_of=wo[f, :cnt[f]], node_of=no[f, :cnt[f]], bow_ids=bi[f, :bc[f]], bow_vals=bv[f, :bc[f]],
                            fv=dict(nodes=fn[f, :nf], start=fs[f, :nf + 1], items=fi[f, :fs[f, nf]])))
        return out
