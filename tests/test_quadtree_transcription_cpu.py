"""A second transcription of ORBextractor::DistributeOctTree / ExtractorNode::DivideNode (src/ORBextractor.cc:480-536, 555-779),
in Python with a list standing in for std::list and the REAL libstdc++ std::sort (through tests/libhostcheck.so) for the
tie-sensitive `sort(..., compareNodes)` -- against the C++ oracle's version on the candidate sets of real frames, every level."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EDGE_THRESHOLD = 19


def _std_sort(pairs):
    L = C.CDLL(os.path.join(HERE, 'libhostcheck.so'))
    n = len(pairs)
    size = (C.c_int * n)(*[p[0] for p in pairs]); ulx = (C.c_int * n)(*[p[1].UL[0] for p in pairs]); out = (C.c_int * n)()
    L.hc_std_sort_order(size, ulx, n, out)
    return [pairs[out[i]] for i in range(n)]


class Node:
    def __init__(self):
        self.UL = self.UR = self.BL = self.BR = (0, 0)
        self.keys = []
        self.noMore = False

    def divide(self, X, Y):
        halfX = int(math.ceil(np.float32(self.UR[0] - self.UL[0]) / 2))
        halfY = int(math.ceil(np.float32(self.BR[1] - self.UL[1]) / 2))
        n1, n2, n3, n4 = Node(), Node(), Node(), Node()
        n1.UL = self.UL; n1.UR = (self.UL[0] + halfX, self.UL[1]); n1.BL = (self.UL[0], self.UL[1] + halfY); n1.BR = (self.UL[0] + halfX, self.UL[1] + halfY)
        n2.UL = n1.UR; n2.UR = self.UR; n2.BL = n1.BR; n2.BR = (self.UR[0], self.UL[1] + halfY)
        n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = self.BL; n3.BR = (n1.BR[0], self.BL[1])
        n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = self.BR
        for k in self.keys:
            if X[k] < n1.UR[0]:
                (n1 if Y[k] < n1.BR[1] else n3).keys.append(k)
            elif Y[k] < n1.BR[1]:
                n2.keys.append(k)
            else:
                n4.keys.append(k)
        for c in (n1, n2, n3, n4):
            if len(c.keys) == 1:
                c.noMore = True
        return n1, n2, n3, n4


def distribute(X, Y, Rsp, minX, maxX, minY, maxY, N):
    nIni = int(math.floor(np.float32(maxX - minX) / np.float32(maxY - minY) + np.float32(0.5)))       # round(), positive
    hX = np.float32(maxX - minX) / np.float32(nIni)
    lNodes, ini = [], []
    for i in range(nIni):
        ni = Node()
        ni.UL = (int(hX * np.float32(i)), 0); ni.UR = (int(hX * np.float32(i + 1)), 0)
        ni.BL = (ni.UL[0], maxY - minY); ni.BR = (ni.UR[0], maxY - minY)
        lNodes.append(ni); ini.append(ni)
    for k in range(len(X)):
        ini[int(X[k] / hX)].keys.append(k)
    kept = []
    for nd in lNodes:
        if len(nd.keys) == 1:
            nd.noMore = True
        if nd.keys:
            kept.append(nd)
    lNodes = kept

    def push_children(node, vec, count):
        added = 0
        for c in node.divide(X, Y):
            if c.keys:
                lNodes.insert(0, c)
                added += 1
                if len(c.keys) > 1:
                    count[0] += 1
                    vec.append((len(c.keys), c))
        return added

    finish = False
    while not finish:
        prevSize = len(lNodes)
        nToExpand = [0]
        vSize = []
        i = 0
        while i < len(lNodes):
            nd = lNodes[i]
            if nd.noMore:
                i += 1
                continue
            i += push_children(nd, vSize, nToExpand)        # children go to the front: the iterator's position shifts
            del lNodes[i]                                   # lit = lNodes.erase(lit)
        if len(lNodes) >= N or len(lNodes) == prevSize:
            finish = True
        elif len(lNodes) + nToExpand[0] * 3 > N:
            while not finish:
                prevSize = len(lNodes)
                vPrev = _std_sort(vSize)
                vSize = []
                for j in range(len(vPrev) - 1, -1, -1):
                    nd = vPrev[j][1]
                    push_children(nd, vSize, [0])
                    lNodes.remove(nd)                       # lNodes.erase(node->lit)
                    if len(lNodes) >= N:
                        break
                if len(lNodes) >= N or len(lNodes) == prevSize:
                    finish = True
    out = []
    for nd in lNodes:
        best = nd.keys[0]
        for k in nd.keys[1:]:
            if Rsp[k] > Rsp[best]:
                best = k
        out.append(best)
    return out


@pytest.mark.parametrize('t,nfeat', [(1, 1000), (6, 1000), (6, 300), (12, 2500)])
def test_distribute_oct_tree_transcription(t, nfeat):
    img = synth.frame(t)
    ex = O.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    ex(img, (0, 1000))
    per_level = ex.tables()['features_per_level']
    checked = 0
    for level in range(8):
        cands = ex.candidates(level)
        if len(cands) == 0:
            continue
        plane = ex.level(level)
        minX = minY = EDGE_THRESHOLD - 3
        maxX, maxY = plane.shape[1] - EDGE_THRESHOLD + 3, plane.shape[0] - EDGE_THRESHOLD + 3
        N = int(per_level[level])
        got = ex.distribute(cands, minX, maxX, minY, maxY, N)
        idx = distribute(cands['x'], cands['y'], cands['response'], minX, maxX, minY, maxY, N)
        assert len(got) == len(idx), (level, len(got), len(idx))
        assert np.array_equal(got['x'], cands['x'][idx]) and np.array_equal(got['y'], cands['y'][idx]) and np.array_equal(got['response'], cands['response'][idx]), level
        checked += 1
    assert checked >= 6
