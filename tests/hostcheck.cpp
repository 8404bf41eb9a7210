// Host-side build of csrc/exact_math.h for CPU tests: the same source the kernels use,
// compiled with g++ -ffp-contract=off, compared against glibc / libstdc++ / the oracle.
#include "../orb_slam3_modified_b200/csrc/exact_math.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

void hc_sincosf_n(const float* a, float* s, float* c, int n) {
    for (int i = 0; i < n; ++i) orbx::sincosf_glibc(a[i], &s[i], &c[i]);
}

// exhaustive sweep over all floats with bit patterns [lo, hi): returns mismatch count vs glibc sincosf
long hc_sincosf_sweep(uint32_t lo, uint32_t hi, uint32_t* first_bad) {
    long bad = 0;
    for (uint32_t u = lo; u < hi; ++u) {
        float y; std::memcpy(&y, &u, 4);
        float s0, c0, s1, c1;
        sincosf(y, &s0, &c0);
        orbx::sincosf_glibc(y, &s1, &c1);
        if (std::memcmp(&s0, &s1, 4) || std::memcmp(&c0, &c1, 4)) { if (!bad && first_bad) *first_bad = u; ++bad; }
    }
    return bad;
}

// exhaustive sweep over all floats with bit patterns [lo, hi): mismatch count of logf_glibc vs the host libm logf
long hc_logf_sweep(uint32_t lo, uint32_t hi, uint32_t* first_bad) {
    long bad = 0;
    for (uint32_t u = lo; u < hi; ++u) {
        float y; std::memcpy(&y, &u, 4);
        const float a = logf(y), b = orbx::logf_glibc(y);
        if (std::memcmp(&a, &b, 4)) { if (!bad && first_bad) *first_bad = u; ++bad; }
    }
    return bad;
}

void hc_fast_atan2_n(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = orbx::fast_atan2_deg(y[i], x[i]);
}

struct SK { int size; int ulx; int id; };
// the permutation REAL libstdc++ std::sort produces for (size, UL.x) keys under the reference's compareNodes
// (src/ORBextractor.cc:538-553: size ascending, then UL.x ascending, everything else a tie)
void hc_std_sort_order(const int* size, const int* ulx, int n, int* order_out) {
    std::vector<SK> b(n);
    for (int i = 0; i < n; ++i) b[i] = {size[i], ulx[i], i};
    std::sort(b.begin(), b.end(), [](const SK& e1, const SK& e2) {
        if (e1.size < e2.size) return true;
        else if (e1.size > e2.size) return false;
        else return e1.ulx < e2.ulx;
    });
    for (int i = 0; i < n; ++i) order_out[i] = b[i].id;
}

// the round-based form of the emulation (one partition step per pending range and round, leaves insertion-sorted independently: what the
// kernel runs with one thread per task) against std::sort; returns 0 if equal
int hc_sort_check_rounds(const int* size, const int* ulx, int n) {
    std::vector<SK> a(n), b(n);
    for (int i = 0; i < n; ++i) { a[i] = {size[i], ulx[i], i}; b[i] = a[i]; }
    auto less = [](const SK& p, const SK& q) { return p.size < q.size || (p.size == q.size && p.ulx < q.ulx); };
    std::vector<orbx::SxTask> q0(n + 2), q1(n + 2), leaf(n + 2);
    orbx::libstdcxx_sort_rounds_host(a.data(), n, less, q0.data(), q1.data(), leaf.data());
    std::sort(b.begin(), b.end(), less);
    for (int i = 0; i < n; ++i) if (a[i].id != b[i].id) return 1;
    return 0;
}

// sorts keys with the emulation and with std::sort using the reference's comparator shape; returns 0 if equal
int hc_sort_check(const int* size, const int* ulx, int n, int* order_out) {
    std::vector<SK> a(n), b(n);
    for (int i = 0; i < n; ++i) { a[i] = {size[i], ulx[i], i}; b[i] = a[i]; }
    auto less = [](const SK& p, const SK& q) { return p.size < q.size || (p.size == q.size && p.ulx < q.ulx); };
    orbx::libstdcxx_sort(a.data(), n, less);
    std::sort(b.begin(), b.end(), less);
    int bad = 0;
    for (int i = 0; i < n; ++i) { if (a[i].id != b[i].id) bad = 1; if (order_out) order_out[i] = a[i].id; }
    return bad;
}
}

// ---------------------------------------------------------------------------
// Sequential model of the list-rebuild formulation used by quadtree_orient_kernel
// (extractor_kernels.cuh): validates the reformulation of DistributeOctTree against the
// oracle's std::list version on the CPU.  Same steps A..H as the kernel, one "thread".
// ---------------------------------------------------------------------------
namespace {
struct QN { int ulx, uly, brx, bry, cnt; };
struct QS { int size, ulx, node; };
inline int qchild(const QN& n, int px, int py, int& mx, int& my) {
    int hx = (int)std::ceil((float)(n.brx - n.ulx) / 2), hy = (int)std::ceil((float)(n.bry - n.uly) / 2);
    mx = n.ulx + hx; my = n.uly + hy;
    return (px < mx ? 0 : 1) + (py < my ? 0 : 2);
}
}

extern "C" int hc_quadtree(const int* xs, const int* ys, const int* resp, int n, int W, int H, int N, int* outIdx) {
    const int nIni = (int)std::round((float)W / H);
    const float hX = (float)W / nIni;
    std::vector<QN> cur;
    std::vector<int> nodeOf(n);
    {
        std::vector<QN> roots(nIni);
        for (int i = 0; i < nIni; ++i) roots[i] = {(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), H, 0};
        for (int i = 0; i < n; ++i) { nodeOf[i] = (int)((float)xs[i] / hX); roots[nodeOf[i]].cnt++; }
        std::vector<int> keep(nIni, -1);
        for (int i = 0; i < nIni; ++i) if (roots[i].cnt > 0) { keep[i] = (int)cur.size(); cur.push_back(roots[i]); }
        for (int i = 0; i < n; ++i) nodeOf[i] = keep[nodeOf[i]];
    }
    int m = (int)cur.size();
    bool finish = n == 0, careful = false;
    std::vector<QS> v;
    while (!finish) {
        const int prevSize = m;
        std::vector<int> proc(m, -1), procNode;
        if (!careful) {
            for (int s = 0; s < m; ++s) if (cur[s].cnt > 1) { proc[s] = (int)procNode.size(); procNode.push_back(s); }
        } else {
            orbx::libstdcxx_sort(v.data(), (int)v.size(), [](const QS& a, const QS& b) { return a.size < b.size || (a.size == b.size && a.ulx < b.ulx); });
            const int q = (int)v.size();
            procNode.resize(q);
            for (int j = 0; j < q; ++j) { proc[v[j].node] = q - 1 - j; procNode[q - 1 - j] = v[j].node; }
        }
        const int P0 = (int)procNode.size();
        std::vector<int> cc(4 * P0, 0);
        for (int i = 0; i < n; ++i) { int pi = proc[nodeOf[i]]; if (pi >= 0) { int mx, my; cc[4 * pi + qchild(cur[nodeOf[i]], xs[i], ys[i], mx, my)]++; } }
        std::vector<int> k(P0), excl(P0 + 1, 0);
        for (int pi = 0; pi < P0; ++pi) { k[pi] = (cc[4*pi]>0)+(cc[4*pi+1]>0)+(cc[4*pi+2]>0)+(cc[4*pi+3]>0); excl[pi + 1] = excl[pi] + k[pi]; }
        int Pn = P0;
        if (careful) for (int pi = 0; pi < P0; ++pi) if (prevSize + excl[pi + 1] - (pi + 1) >= N) { Pn = pi + 1; break; }
        for (int pi = Pn; pi < P0; ++pi) proc[procNode[pi]] = -1;
        const int totalChildren = excl[Pn];
        int nKeep = 0;
        for (int s = 0; s < m; ++s) nKeep += proc[s] < 0;
        std::vector<QN> nxt(totalChildren + nKeep);
        std::vector<int> childPos(4 * P0, -1), keepPos(m, -1);
        for (int pi = 0; pi < Pn; ++pi) {
            const int blockStart = totalChildren - excl[pi + 1];
            const QN pn = cur[procNode[pi]];
            int mx, my; qchild(pn, 0, 0, mx, my);
            int after = 0;
            for (int ch = 3; ch >= 0; --ch) if (cc[4 * pi + ch] > 0) {
                const int pos = blockStart + after++;
                childPos[4 * pi + ch] = pos;
                nxt[pos] = {(ch & 1) ? mx : pn.ulx, (ch & 2) ? my : pn.uly, (ch & 1) ? pn.brx : mx, (ch & 2) ? pn.bry : my, cc[4 * pi + ch]};
            }
        }
        { int r = 0; for (int s = 0; s < m; ++s) if (proc[s] < 0) { keepPos[s] = totalChildren + r++; nxt[keepPos[s]] = cur[s]; } }
        v.clear();
        for (int i = 0; i < 4 * Pn; ++i) if (cc[i] > 1) v.push_back({cc[i], nxt[childPos[i]].ulx, childPos[i]});
        for (int i = 0; i < n; ++i) {
            int s = nodeOf[i], pi = proc[s];
            if (pi >= 0) { int mx, my; nodeOf[i] = childPos[4 * pi + qchild(cur[s], xs[i], ys[i], mx, my)]; }
            else nodeOf[i] = keepPos[s];
        }
        cur.swap(nxt);
        m = (int)cur.size();
        if (m >= N || m == prevSize) finish = true;
        else if (!careful && m + 3 * (int)v.size() > N) careful = true;
    }
    std::vector<long long> best(m, -1);
    for (int i = 0; i < n; ++i) {
        long long key = ((long long)resp[i] << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
        if (key > best[nodeOf[i]]) best[nodeOf[i]] = key;
    }
    for (int s = 0; s < m; ++s) outIdx[s] = (int)(0xFFFFFFFFu - (unsigned)(best[s] & 0xFFFFFFFFll));
    return m;
}
