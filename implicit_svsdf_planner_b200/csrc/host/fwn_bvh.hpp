// host/fwn_bvh.hpp — the mesh functor's winding-number hierarchy, built on the host.
//
// BasicShape::getonlySDF_igl (utils/include/utils/Shape.hpp:332-340) takes its sign from
// igl::fast_winding_number(fwn_bvh, 2.0, p) (igl/fast_winding_number.cpp:439-457), i.e. from the HDK's
// UT_SolidAngle<float,float>: a 4-way bounding volume hierarchy over the triangles (UT_BVH<4>::init, surface-area heuristic)
// whose nodes carry an order-2 Taylor expansion of the solid angle of everything below them; a query descends only into
// children closer than accuracy_scale x their radius and sums the expansions of the others — all in SINGLE precision.
// The value is therefore an approximation (2e-3 off the exact winding number on the reference's own meshes), and to agree
// with the reference to better than that one has to reproduce the approximation itself: the same tree, the same
// coefficients, the same traversal.  This file is a re-implementation of that algorithm from its definition
// (igl/FastWindingNumberForSoups.h: BVH<N>::init / initNode / multiSplit / split / nthElement / partitionByCentre
// :4580-5990, UT_SolidAngle::init with its PrecomputeFunctors :6544-7135, computeSolidAngle :7149-7284,
// UTsignedSolidAngleTri :6071-6110), written against plain arrays for upload to the device; tests compare the tree
// node by node, the coefficients and the winding numbers with the reference's own code compiled where it lies
// (oracle/_ref/libref_fwn.so, tests/test_oracle_mesh.py).  Float arithmetic follows the reference's operation order
// (vector / scalar is a multiplication by the reciprocal there, sums run left to right); the host compiler must not
// contract a * b + c (x86-64 baseline does not).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace svsdf {
namespace host {

struct FwnBvh {
    static constexpr uint32_t EMPTY = 0xffffffffu, INTERNAL = 0x80000000u;
    static constexpr int kDataRows = 23;  // floats-per-child rows of the per-node record (see build_data)
    int nn = 0, nv = 0, nf = 0;
    std::vector<uint32_t> child;  // [nn][4]: triangle index, INTERNAL | node index, or EMPTY
    std::vector<float> data;      // [nn][23][4]: per child (lane) — maxPDist2, P[3], N[3], NijDiag[3], Nxy+Nyx, Nyz+Nzy, Nzx+Nxz,
                                  //   NijkDiag[3], SumPermuteNxyz, 2Nxxy+Nyxx, 2Nxxz+Nzxx, 2Nyyz+Nzyy, 2Nyyx+Nxyy, 2Nzzx+Nxzz, 2Nzzy+Nyzz
    std::vector<float> cbox;      // [nn][4][6]: child bounding boxes (min xyz, max xyz); not part of the reference record, used
                                  //   by the device closest-triangle search
    std::vector<float> U;         // nv x 3 positions in float (fast_winding_number.cpp:393-398)
    std::vector<int> F;           // nf x 3

    // ---- boxes ---------------------------------------------------------------------------------------------------------
    struct Box {
        float v[3][2];
        void init_empty() { for (int a = 0; a < 3; ++a) { v[a][0] = std::numeric_limits<float>::max(); v[a][1] = -std::numeric_limits<float>::max(); } }
        void combine(const Box &s) {
            for (int a = 0; a < 3; ++a) {
                v[a][0] = (v[a][0] < s.v[a][0]) ? v[a][0] : s.v[a][0];
                v[a][1] = (v[a][1] > s.v[a][1]) ? v[a][1] : s.v[a][1];
            }
        }
        float half_area() const {
            const float d0 = v[0][1] - v[0][0], d1 = v[1][1] - v[1][0], d2 = v[2][1] - v[2][0];
            return d0 * d1 + d1 * d2 + d2 * d0;
        }
        float centre2(int axis) const { return v[axis][0] + v[axis][1]; }
    };

    // ---- tree construction (UT_BVH<4>::init<BOX_AREA>) ----------------------------------------------------------------------
    std::vector<Box> tb;  // triangle boxes
    struct Node { uint32_t c[4]; };
    std::vector<Node> nodes;

    static void partition_by_centre(const Box *boxes, uint32_t *ind, uint32_t *end, int axis, float pivot, uint32_t *&ps, uint32_t *&pe) {
        uint32_t *pivot_start = ind, *pivot_end = ind;
        for (uint32_t *p = ind; p != end; ++p) {
            const float s = boxes[*p].centre2(axis);
            if (s < pivot) {
                if (p != pivot_start) {
                    if (pivot_start == pivot_end) { const uint32_t t = *p; *p = *pivot_start; *pivot_start = t; }
                    else { const uint32_t t = *p; *p = *pivot_end; *pivot_end = *pivot_start; *pivot_start = t; }
                }
                ++pivot_start; ++pivot_end;
            } else if (s == pivot) {
                if (p != pivot_end) { const uint32_t t = *p; *p = *pivot_end; *pivot_end = t; }
                ++pivot_end;
            }
        }
        ps = pivot_start; pe = pivot_end;
    }
    static void nth_element(const Box *boxes, uint32_t *ind, uint32_t *end, int axis, uint32_t *nth) {
        for (;;) {
            float pv[3] = {boxes[ind[0]].centre2(axis), boxes[ind[(end - ind) / 2]].centre2(axis), boxes[*(end - 1)].centre2(axis)};
            if (pv[0] < pv[1]) std::swap(pv[0], pv[1]);
            if (pv[0] < pv[2]) std::swap(pv[0], pv[2]);
            if (pv[1] < pv[2]) std::swap(pv[1], pv[2]);
            uint32_t *ps, *pe;
            partition_by_centre(boxes, ind, end, axis, pv[1], ps, pe);
            if (nth < ps) end = ps;
            else if (nth < pe) return;
            else ind = pe;
            if (end <= ind + 1) return;
        }
    }

    static void split(const Box &mm, const Box *boxes, uint32_t *ind, uint32_t n, uint32_t *&split_ind, Box *sb) {
        if (n < 2) { sb[0] = mm; sb[1] = mm; split_ind = ind + n; return; }  // not reached: a node is split only when it holds >= 2 items
        if (n == 2) { sb[0] = boxes[ind[0]]; sb[1] = boxes[ind[1]]; split_ind = ind + 1; return; }
        constexpr uint32_t SMALL = 6, NSPANS = 16, NSPLITS = 15, MID = 32, MINFRAC = 16;
        if (n <= SMALL) {  // all (2^(n-1)) - 1 partitions with box 0 in part 0
            Box lb[SMALL];
            for (uint32_t b = 0; b < n; ++b) lb[b] = boxes[ind[b]];
            const uint32_t limit = 1u << (n - 1);
            uint32_t best = EMPTY;
            float best_h = 0.0f;
            for (uint32_t bits = 1; bits < limit; ++bits) {
                Box s[2];
                s[0] = lb[0];
                s[1].init_empty();
                uint32_t cnt[2] = {1, 0};
                for (uint32_t bit = 0; bit < n - 1; ++bit) {
                    const uint32_t dest = (bits >> bit) & 1u;
                    s[dest].combine(lb[bit + 1]);
                    ++cnt[dest];
                }
                const float h = s[0].half_area() * cnt[0] + s[1].half_area() * cnt[1];
                if (best == EMPTY || h < best_h) { best = bits; best_h = h; sb[0] = s[0]; sb[1] = s[1]; }
            }
            uint32_t li[SMALL - 1];
            for (uint32_t b = 0; b < n - 1; ++b) li[b] = ind[b + 1];
            uint32_t *dst = ind + 1;
            for (uint32_t bit = 0; bit < n - 1; ++bit)
                if (!((best >> bit) & 1u)) *dst++ = li[bit];
            split_ind = dst;
            for (uint32_t bit = 0; bit < n - 1; ++bit)
                if ((best >> bit) & 1u) *dst++ = li[bit];
            return;
        }
        int axis = 0;
        float len = mm.v[0][1] - mm.v[0][0];
        for (int a = 1; a < 3; ++a) {
            const float l = mm.v[a][1] - mm.v[a][0];
            if (l > len) { axis = a; len = l; }
        }
        if (!(len > 0.0f)) { split_ind = ind + n / 2; sb[0] = mm; sb[1] = mm; return; }
        if (n <= MID) {  // sort by centre along the axis (stable), try every split
            float mid2[MID];
            for (uint32_t i = 0; i < n; ++i) mid2[i] = boxes[ind[i]].centre2(axis);
            uint32_t li[MID];
            for (uint32_t i = 0; i < n; ++i) li[i] = i;
            // (insertion sort of four chunks + three stable merges in the reference = one stable sort)
            std::stable_sort(li, li + n, [&mid2](uint32_t a, uint32_t b) { return mid2[a] < mid2[b]; });
            for (uint32_t i = 0; i < n; ++i) li[i] = ind[li[i]];
            for (uint32_t i = 0; i < n; ++i) ind[i] = li[i];
            Box left[MID - 1], right[MID - 1];
            const uint32_t ns = n - 1;
            Box acc = boxes[li[0]];
            left[0] = acc;
            for (uint32_t i = 1; i < ns; ++i) { acc.combine(boxes[li[i]]); left[i] = acc; }
            acc = boxes[li[ns - 1]];   // (sic: the reference seeds the right accumulation with element ns - 1, then adds ns - 1 .. 1)
            right[ns - 1] = acc;
            for (uint32_t i = ns - 1; i > 0; --i) { acc.combine(boxes[li[i]]); right[i - 1] = acc; }
            uint32_t bs = 0;
            float bh = left[0].half_area() + right[0].half_area() * (n - 1);
            for (uint32_t s = 1; s < ns; ++s) {
                const float h = left[s].half_area() * (s + 1) + right[s].half_area() * (n - (s + 1));
                if (h < bh) { bs = s; bh = h; }
            }
            split_ind = ind + bs + 1;
            sb[0] = left[bs];
            sb[1] = right[bs];
            return;
        }
        const float axis_min = mm.v[axis][0];
        Box span[NSPANS];
        uint32_t cnt[NSPANS];
        for (uint32_t i = 0; i < NSPANS; ++i) { span[i].init_empty(); cnt[i] = 0; }
        const float axis_min_x2 = 2 * axis_min;
        const float scale = (float(1.0 / 2) * NSPANS) / len;
        for (uint32_t i = 0; i < n; ++i) {
            const Box &b = boxes[ind[i]];
            const float sum = b.centre2(axis);
            int si = int((sum - axis_min_x2) * scale);
            si = si < 0 ? 0 : (si > int(NSPANS - 1) ? int(NSPANS - 1) : si);
            ++cnt[si];
            span[si].combine(b);
        }
        Box left[NSPLITS], right[NSPLITS];
        Box acc = span[0];
        left[0] = acc;
        for (uint32_t i = 1; i < NSPLITS; ++i) { acc.combine(span[i]); left[i] = acc; }
        acc = span[NSPANS - 1];
        right[NSPLITS - 1] = acc;
        for (uint32_t i = NSPLITS - 1; i > 0; --i) { acc.combine(span[i]); right[i - 1] = acc; }
        uint32_t lc[NSPLITS];
        uint32_t ca = cnt[0];
        lc[0] = ca;
        for (uint32_t i = 1; i < NSPLITS; ++i) { ca += cnt[i]; lc[i] = ca; }
        const uint32_t min_count = n / MINFRAC, max_count = (uint32_t)(((MINFRAC - 1) * uint64_t(n)) / MINFRAC);
        float smallest = std::numeric_limits<float>::infinity();
        int si = -1;
        for (uint32_t s = 0; s < NSPLITS; ++s) {
            const uint32_t l = lc[s];
            if (l < min_count || l > max_count) continue;
            const uint32_t r = n - l;
            const float h = l * left[s].half_area() + r * right[s].half_area();
            if (h < smallest) { smallest = h; si = (int)s; }
        }
        uint32_t *const end = ind + n;
        if (si == -1) {
            uint32_t *nth;
            if (lc[0] > max_count) nth = ind + max_count;
            else if (lc[NSPLITS - 1] < min_count) nth = ind + min_count;
            else nth = ind + n / 2;
            nth_element(boxes, ind, end, axis, nth);
            split_ind = nth;
            Box lb = boxes[ind[0]];
            for (uint32_t *p = ind + 1; p < nth; ++p) lb.combine(boxes[*p]);
            Box rb = boxes[nth[0]];
            for (uint32_t *p = nth + 1; p < end; ++p) rb.combine(boxes[*p]);
            sb[0] = lb; sb[1] = rb;
            return;
        }
        const float pivot = axis_min_x2 + (si + 1) * len / (NSPANS / 2);
        uint32_t *ps, *pe;
        partition_by_centre(boxes, ind, end, axis, pivot, ps, pe);
        split_ind = ind + lc[si];
        if (split_ind >= ps && split_ind <= pe) { sb[0] = left[si]; sb[1] = right[si]; return; }
        if (split_ind < ps) split_ind = ps;
        else split_ind = pe;
        if (split_ind == ind) ++split_ind;
        else if (split_ind == end) --split_ind;
        Box lb = boxes[ind[0]];
        for (uint32_t *p = ind + 1; p < split_ind; ++p) lb.combine(boxes[*p]);
        Box rb = boxes[split_ind[0]];
        for (uint32_t *p = split_ind + 1; p < end; ++p) rb.combine(boxes[*p]);
        sb[0] = lb; sb[1] = rb;
    }

    static void multi_split(const Box &mm, const Box *boxes, uint32_t *ind, uint32_t n, uint32_t *sub[5], Box sb[4]) {
        sub[0] = ind;
        sub[2] = ind + n;
        split(mm, boxes, ind, n, sub[1], &sb[0]);
        float area[4];
        area[0] = sb[0].half_area();
        area[1] = sb[1].half_area();
        for (uint32_t nsub = 2; nsub < 4; ++nsub) {
            uint32_t choice = EMPTY;
            float maxh = 0.0f;
            for (uint32_t i = 0; i < nsub; ++i) {
                const uint32_t cnt = (uint32_t)(sub[i + 1] - sub[i]);
                if (cnt > 1) {
                    const float h = area[i] * cnt;
                    if (choice == EMPTY || h > maxh) { choice = i; maxh = h; }
                }
            }
            uint32_t *s0 = sub[choice], *s1 = sub[choice + 1];
            for (uint32_t i = nsub; i > choice; --i) sub[i + 1] = sub[i];
            for (uint32_t i = nsub - 1; i > choice; --i) sb[i + 1] = sb[i];
            for (uint32_t i = nsub - 1; i > choice; --i) area[i + 1] = area[i];
            const Box parent = sb[choice];
            split(parent, boxes, s0, (uint32_t)(s1 - s0), sub[choice + 1], &sb[choice]);
            area[choice] = sb[choice].half_area();
            area[choice + 1] = sb[choice + 1].half_area();
        }
    }

    // nodes of a subtree follow their parent's child slots in the order the reference's serial recursion appends them
    void init_node(uint32_t node_index, const Box &mm, uint32_t *ind, uint32_t n) {
        if (n <= 4) {
            for (uint32_t i = 0; i < n; ++i) nodes[node_index].c[i] = ind[i];
            for (uint32_t i = n; i < 4; ++i) nodes[node_index].c[i] = EMPTY;
            return;
        }
        uint32_t *sub[5];
        Box sb[4];
        multi_split(mm, tb.data(), ind, n, sub, sb);
        for (uint32_t i = 0; i < 4; ++i)
            if (sub[i + 1] - sub[i] == 1) nodes[node_index].c[i] = sub[i][0];
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t sn = (uint32_t)(sub[i + 1] - sub[i]);
            if (sn != 1) {
                const uint32_t start = (uint32_t)nodes.size();
                nodes[node_index].c[i] = start | INTERNAL;
                nodes.push_back(Node());
                init_node(start, sb[i], sub[i], sn);
            }
        }
    }

    // ---- per-node expansion data (UT_SolidAngle::init, PrecomputeFunctors) -----------------------------------------------
    struct Local {
        Box box;
        float P[3], areaP[3], N[3], area;
        float Nii[3], Nxy, Nyx, Nyz, Nzy, Nzx, Nxz;
        float Niii[3], sumperm, xxy, xxz, yyz, yyx, zzx, zzy;
    };
    float *row(int node, int r) { return &data[((size_t)node * kDataRows + r) * 4]; }
    const float *row(int node, int r) const { return &data[((size_t)node * kDataRows + r) * 4]; }

    static void integrals(const float a[3], const float b[3], const float c[3], const float P[3], float *ii, float *ij, float *ik, int i) {
        float oab[3], oac[3], ocb[3];
        for (int d = 0; d < 3; ++d) { oab[d] = b[d] - a[d]; oac[d] = c[d] - a[d]; ocb[d] = b[d] - c[d]; }
        const float t = oab[i] / oac[i];
        const int j = (i == 2) ? 0 : i + 1, k = (j == 2) ? 0 : j + 1;
        const float jdiff = t * oac[j] - oab[j], kdiff = t * oac[k] - oab[k];
        float ca[3], cc[3];
        ca[0] = (jdiff * oab[k] - kdiff * oab[j]); ca[1] = kdiff * oab[i]; ca[2] = jdiff * oab[i];
        cc[0] = (jdiff * ocb[k] - kdiff * ocb[j]); cc[1] = kdiff * ocb[i]; cc[2] = jdiff * ocb[i];
        const float sa = std::sqrt(ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2]);
        const float sc = std::sqrt(cc[0] * cc[0] + cc[1] * cc[1] + cc[2] * cc[2]);
        const float Pai = a[i] - P[i], Pci = c[i] - P[i];
        const float ia = sa * (float(0.5) * Pai * Pai + float(2.0 / 3.0) * Pai * oab[i] + float(0.25) * oab[i] * oab[i]);
        const float ic = sc * (float(0.5) * Pci * Pci + float(2.0 / 3.0) * Pci * ocb[i] + float(0.25) * ocb[i] * ocb[i]);
        *ii = ia + ic;
        int jk = j;
        float *integral = ij;
        float diff = jdiff;
        for (;;) {
            if (integral) {
                const float obmid = b[jk] + float(0.5) * diff;
                const float oabmid = obmid - a[jk], ocbmid = obmid - c[jk];
                const float Paj = a[jk] - P[jk], Pcj = c[jk] - P[jk];
                const float xa = sa * (float(0.5) * Pai * Paj + float(1.0 / 3.0) * Pai * oabmid + float(1.0 / 3.0) * Paj * oab[i] + float(0.25) * oab[i] * oabmid);
                const float xc = sc * (float(0.5) * Pci * Pcj + float(1.0 / 3.0) * Pci * ocbmid + float(1.0 / 3.0) * Pcj * ocb[i] + float(0.25) * ocb[i] * ocbmid);
                *integral = xa + xc;
            }
            if (jk == k) break;
            jk = k;
            integral = ik;
            diff = kdiff;
        }
    }

    void item(int tri, Local &L) const {
        const float *a = &U[3 * (size_t)F[3 * tri]], *b = &U[3 * (size_t)F[3 * tri + 1]], *c = &U[3 * (size_t)F[3 * tri + 2]];
        float ab[3], ac[3];
        for (int d = 0; d < 3; ++d) { ab[d] = b[d] - a[d]; ac[d] = c[d] - a[d]; }
        L.box = tb[tri];
        float N[3];
        N[0] = float(0.5) * (ab[1] * ac[2] - ab[2] * ac[1]);
        N[1] = float(0.5) * (ab[2] * ac[0] - ab[0] * ac[2]);
        N[2] = float(0.5) * (ab[0] * ac[1] - ab[1] * ac[0]);
        float area2 = N[0] * N[0];
        area2 += N[1] * N[1];
        area2 += N[2] * N[2];
        const float area = std::sqrt(area2);
        const float third = 1 / float(3);   // (a + b + c) / 3: a multiplication by the reciprocal
        float P[3];
        for (int d = 0; d < 3; ++d) P[d] = ((a[d] + b[d]) + c[d]) * third;
        for (int d = 0; d < 3; ++d) { L.P[d] = P[d]; L.areaP[d] = P[d] * area; L.N[d] = N[d]; }
        L.area = area;
        L.Nii[0] = L.Nii[1] = L.Nii[2] = 0;
        L.Nxy = L.Nyx = L.Nyz = L.Nzy = L.Nzx = L.Nxz = 0;
        if (area == 0) {
            L.Niii[0] = L.Niii[1] = L.Niii[2] = 0;
            L.sumperm = L.xxy = L.xxz = L.yyz = L.yyx = L.zzx = L.zzy = 0;
            return;
        }
        const float inv_area = 1 / area;
        const float n[3] = {N[0] * inv_area, N[1] * inv_area, N[2] * inv_area};
        const float *val[3] = {a, b, c};
        int ord[3][3];
        float dd[3];
        for (int ax = 0; ax < 3; ++ax) {
            int *o = ord[ax];
            o[0] = 0; o[1] = 1; o[2] = 2;
            if (a[ax] > b[ax]) std::swap(o[0], o[1]);
            if (val[o[0]][ax] > c[ax]) std::swap(o[0], o[2]);
            if (val[o[1]][ax] > val[o[2]][ax]) std::swap(o[1], o[2]);
            dd[ax] = val[o[2]][ax] - val[o[0]][ax];
        }
        const float dx = dd[0], dy = dd[1], dz = dd[2];
        float ixx = 0, ixy = 0, iyy = 0, iyz = 0, izz = 0, izx = 0;
        if (dx > 0) integrals(val[ord[0][0]], val[ord[0][1]], val[ord[0][2]], P, &ixx, ((dx >= dy && dy > 0) ? &ixy : nullptr), ((dx >= dz && dz > 0) ? &izx : nullptr), 0);
        if (dy > 0) integrals(val[ord[1][0]], val[ord[1][1]], val[ord[1][2]], P, &iyy, ((dy >= dz && dz > 0) ? &iyz : nullptr), ((dx < dy && dx > 0) ? &ixy : nullptr), 1);
        if (dz > 0) integrals(val[ord[2][0]], val[ord[2][1]], val[ord[2][2]], P, &izz, ((dx < dz && dx > 0) ? &izx : nullptr), ((dy < dz && dy > 0) ? &iyz : nullptr), 2);
        L.Niii[0] = ixx * n[0]; L.Niii[1] = iyy * n[1]; L.Niii[2] = izz * n[2];
        L.sumperm = 2 * (n[0] * iyz + n[1] * izx + n[2] * ixy);
        const float Nxxy = n[0] * ixy, Nxxz = n[0] * izx, Nyyz = n[1] * iyz, Nyyx = n[1] * ixy, Nzzx = n[2] * izx, Nzzy = n[2] * iyz;
        L.xxy = 2 * Nxxy + n[1] * ixx;
        L.xxz = 2 * Nxxz + n[2] * ixx;
        L.yyz = 2 * Nyyz + n[2] * iyy;
        L.yyx = 2 * Nyyx + n[0] * iyy;
        L.zzx = 2 * Nzzx + n[0] * izz;
        L.zzy = 2 * Nzzy + n[1] * izz;
    }

    void post(int node, Local &out, int nch, const Local *ch) {
        float N[3] = {ch[0].N[0], ch[0].N[1], ch[0].N[2]};
        float areaP[3] = {ch[0].areaP[0], ch[0].areaP[1], ch[0].areaP[2]};
        float area = ch[0].area;
        for (int i = 0; i < 4; ++i)
            for (int r = 0; r < kDataRows; ++r) row(node, r)[i] = 0.0f;
        for (int i = 0; i < nch; ++i) {
            for (int d = 0; d < 3; ++d) { row(node, 4 + d)[i] = ch[i].N[d]; row(node, 1 + d)[i] = ch[i].P[d]; }
            if (i > 0) {
                for (int d = 0; d < 3; ++d) { N[d] += ch[i].N[d]; areaP[d] += ch[i].areaP[d]; }
                area += ch[i].area;
            }
        }
        for (int d = 0; d < 3; ++d) { out.N[d] = N[d]; out.areaP[d] = areaP[d]; }
        out.area = area;
        Box box = ch[0].box;
        for (int i = 1; i < nch; ++i) box.combine(ch[i].box);
        float avg[3];
        if (area > 0) {
            const float inv = 1 / area;
            for (int d = 0; d < 3; ++d) avg[d] = areaP[d] * inv;
        } else {
            for (int d = 0; d < 3; ++d) avg[d] = float(0.5) * (box.v[d][0] + box.v[d][1]);
        }
        for (int d = 0; d < 3; ++d) out.P[d] = avg[d];
        out.box = box;
        for (int i = 0; i < nch; ++i) {
            float m2 = 0.0f;
            for (int d = 0; d < 3; ++d) {
                const float lo = ch[i].P[d] - ch[i].box.v[d][0], hi = ch[i].box.v[d][1] - ch[i].P[d];
                const float m = lo > hi ? lo : hi;   // SYSmax(a, b) = (a > b) ? a : b
                if (d == 0) m2 = m * m;
                else m2 += m * m;
            }
            row(node, 0)[i] = m2;
            for (int d = 0; d < 3; ++d) { cbox[((size_t)node * 4 + i) * 6 + d] = ch[i].box.v[d][0]; cbox[((size_t)node * 4 + i) * 6 + 3 + d] = ch[i].box.v[d][1]; }
        }
        for (int i = nch; i < 4; ++i) row(node, 0)[i] = std::numeric_limits<float>::infinity();
        // order 2
        for (int d = 0; d < 3; ++d) { out.Nii[d] = ch[0].Nii[d]; out.Niii[d] = ch[0].Niii[d]; }
        out.Nxy = out.Nyx = out.Nyz = out.Nzy = out.Nzx = out.Nxz = 0;
        out.sumperm = ch[0].sumperm; out.xxy = ch[0].xxy; out.xxz = ch[0].xxz; out.yyz = ch[0].yyz; out.yyx = ch[0].yyx;
        out.zzx = ch[0].zzx; out.zzy = ch[0].zzy;
        for (int i = 1; i < nch; ++i) {
            for (int d = 0; d < 3; ++d) { out.Nii[d] += ch[i].Nii[d]; out.Niii[d] += ch[i].Niii[d]; }
            out.sumperm += ch[i].sumperm; out.xxy += ch[i].xxy; out.xxz += ch[i].xxz; out.yyz += ch[i].yyz; out.yyx += ch[i].yyx;
            out.zzx += ch[i].zzx; out.zzy += ch[i].zzy;
        }
        for (int i = 0; i < nch; ++i) {
            for (int d = 0; d < 3; ++d) { row(node, 7 + d)[i] = ch[i].Nii[d]; row(node, 13 + d)[i] = ch[i].Niii[d]; }
            row(node, 10)[i] = ch[i].Nxy + ch[i].Nyx;
            row(node, 11)[i] = ch[i].Nyz + ch[i].Nzy;
            row(node, 12)[i] = ch[i].Nzx + ch[i].Nxz;
            row(node, 16)[i] = ch[i].sumperm;
            row(node, 17)[i] = ch[i].xxy; row(node, 18)[i] = ch[i].xxz; row(node, 19)[i] = ch[i].yyz; row(node, 20)[i] = ch[i].yyx;
            row(node, 21)[i] = ch[i].zzx; row(node, 22)[i] = ch[i].zzy;
        }
        for (int i = 0; i < nch; ++i) {
            const Local &c = ch[i];
            const float d[3] = {c.P[0] - out.P[0], c.P[1] - out.P[1], c.P[2] - out.P[2]};
            const float *Nc = c.N;
            for (int k = 0; k < 3; ++k) out.Nii[k] += Nc[k] * d[k];
            const float Nxy = c.Nxy + Nc[0] * d[1], Nyx = c.Nyx + Nc[1] * d[0], Nyz = c.Nyz + Nc[1] * d[2];
            const float Nzy = c.Nzy + Nc[2] * d[1], Nzx = c.Nzx + Nc[2] * d[0], Nxz = c.Nxz + Nc[0] * d[2];
            out.Nxy += Nxy; out.Nyx += Nyx; out.Nyz += Nyz; out.Nzy += Nzy; out.Nzx += Nzx; out.Nxz += Nxz;
            for (int k = 0; k < 3; ++k) out.Niii[k] += (float(2) * d[k]) * c.Nii[k] + (d[k] * d[k]) * Nc[k];
            out.sumperm += (d[0] * (Nyz + Nzy) + d[1] * (Nzx + Nxz) + d[2] * (Nxy + Nyx));
            out.xxy += 2 * (d[1] * c.Nii[0] + d[0] * c.Nxy + Nc[0] * d[0] * d[1]) + 2 * c.Nyx * d[0] + Nc[1] * d[0] * d[0];
            out.xxz += 2 * (d[2] * c.Nii[0] + d[0] * c.Nxz + Nc[0] * d[0] * d[2]) + 2 * c.Nzx * d[0] + Nc[2] * d[0] * d[0];
            out.yyz += 2 * (d[2] * c.Nii[1] + d[1] * c.Nyz + Nc[1] * d[1] * d[2]) + 2 * c.Nzy * d[1] + Nc[2] * d[1] * d[1];
            out.yyx += 2 * (d[0] * c.Nii[1] + d[1] * c.Nyx + Nc[1] * d[1] * d[0]) + 2 * c.Nxy * d[1] + Nc[0] * d[1] * d[1];
            out.zzx += 2 * (d[0] * c.Nii[2] + d[2] * c.Nzx + Nc[2] * d[2] * d[0]) + 2 * c.Nxz * d[2] + Nc[0] * d[2] * d[2];
            out.zzy += 2 * (d[1] * c.Nii[2] + d[2] * c.Nzy + Nc[2] * d[2] * d[1]) + 2 * c.Nyz * d[2] + Nc[1] * d[2] * d[2];
        }
    }

    void precompute(int node, Local &out) {
        Local ch[4];
        int s;
        for (s = 0; s < 4; ++s) {
            const uint32_t c = nodes[node].c[s];
            if (c & INTERNAL) {
                if (c == EMPTY) break;
                precompute((int)(c & ~INTERNAL), ch[s]);
            } else {
                item((int)c, ch[s]);
            }
        }
        post(node, out, s, ch);
    }

    // V: nv x 3 doubles (already in the shape frame, as BasicShape's constructor leaves them), Fi: nf x 3
    void build(const double *V, int nv_, const int *Fi, int nf_) {
        nv = nv_; nf = nf_;
        U.resize((size_t)nv * 3);
        for (size_t i = 0; i < (size_t)nv * 3; ++i) U[i] = (float)V[i];
        F.assign(Fi, Fi + (size_t)nf * 3);
        tb.resize(nf);
        for (int t = 0; t < nf; ++t) {
            Box &b = tb[t];
            for (int d = 0; d < 3; ++d) b.v[d][0] = b.v[d][1] = U[3 * (size_t)F[3 * t] + d];
            for (int k = 1; k < 3; ++k)
                for (int d = 0; d < 3; ++d) {
                    const float p = U[3 * (size_t)F[3 * t + k] + d];
                    b.v[d][0] = b.v[d][0] < p ? b.v[d][0] : p;   // SYSmin / SYSmax
                    b.v[d][1] = b.v[d][1] > p ? b.v[d][1] : p;
                }
        }
        nodes.clear();
        nn = 0;
        child.clear(); data.clear(); cbox.clear();
        if (nf == 0) return;
        std::vector<uint32_t> ind(nf);
        for (int i = 0; i < nf; ++i) ind[i] = (uint32_t)i;
        // (boxes with NaN / Inf corners are excluded by the reference; meshes are validated before they get here)
        Box mm = tb[ind[0]];
        for (int i = 1; i < nf; ++i) mm.combine(tb[ind[i]]);
        nodes.reserve(nf);
        nodes.push_back(Node());
        init_node(0, mm, ind.data(), (uint32_t)nf);
        nn = (int)nodes.size();
        child.resize((size_t)nn * 4);
        for (int i = 0; i < nn; ++i)
            for (int c = 0; c < 4; ++c) child[4 * (size_t)i + c] = nodes[i].c[c];
        data.assign((size_t)nn * kDataRows * 4, 0.0f);
        cbox.assign((size_t)nn * 4 * 6, 0.0f);
        Local root;
        precompute(0, root);
    }

    // ---- evaluation on the host (validation of the hierarchy; the kernels carry their own copy of this traversal) ------
    // atan2f as glibc computes it (fdlibm e_atan2f.c / s_atanf.c, "huge" threshold 2^25): only IEEE float operations, so
    // the device copy (csrc/svsdf_sincos.cuh: atan2f_portable) returns the same bits; tests compare it with the C library's
    // atan2f bit for bit on 1e7 arguments.
    static inline int32_t fword(float x) { int32_t i; std::memcpy(&i, &x, 4); return i; }
    static float atanf_portable(float x) {
        const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
        const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
        const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                              6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
        const int32_t hx = fword(x), ix = hx & 0x7fffffff;
        int id;
        if (ix >= 0x4c000000) {  // |x| >= 2^25
            if (ix > 0x7f800000) return x + x;
            return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
        }
        if (ix < 0x3ee00000) {  // |x| < 0.4375
            if (ix < 0x31000000) return x;  // |x| < 2^-29
            id = -1;
        } else {
            x = std::fabs(x);
            if (ix < 0x3f980000) {
                if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
                else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
            } else {
                if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
                else { id = 3; x = -1.0f / x; }
            }
        }
        const float z = x * x, w = z * z;
        const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
        const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
        if (id < 0) return x - x * (s1 + s2);
        const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
        return hx < 0 ? -r : r;
    }
    static float atan2f_portable(float y, float x) {
        const float pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
        const int32_t hx = fword(x), ix = hx & 0x7fffffff, hy = fword(y), iy = hy & 0x7fffffff;
        if (ix >= 0x7f800000 || iy >= 0x7f800000) return std::atan2(y, x);  // inf / nan: not reached by the solid-angle code
        if (hx == 0x3f800000) return atanf_portable(y);
        const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
        if (iy == 0) {
            switch (m) {
                case 0:
                case 1: return y;
                case 2: return pi;
                default: return -pi;
            }
        }
        if (ix == 0) return hy < 0 ? -pi_o_2 : pi_o_2;
        const int k = (iy - ix) >> 23;
        float z;
        if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
        else if (hx < 0 && k < -60) z = 0.0f;
        else z = atanf_portable(std::fabs(y / x));
        switch (m) {
            case 0: return z;
            case 1: return -z;
            case 2: return pi - (z - pi_lo);
            default: return (z - pi_lo) - pi;
        }
    }
    static float tri_solid_angle(const float *a, const float *b, const float *c, const float *q) {
        float qa[3], qb[3], qc[3];
        for (int d = 0; d < 3; ++d) { qa[d] = a[d] - q[d]; qb[d] = b[d] - q[d]; qc[d] = c[d] - q[d]; }
        auto len = [](const float *v) { float r = v[0] * v[0]; r += v[1] * v[1]; r += v[2] * v[2]; return std::sqrt(r); };
        const float al = len(qa), bl = len(qb), cl = len(qc);
        if (al == 0 || bl == 0 || cl == 0) return 0.0f;
        const float ia = 1 / al, ib = 1 / bl, ic = 1 / cl;
        for (int d = 0; d < 3; ++d) { qa[d] *= ia; qb[d] *= ib; qc[d] *= ic; }
        const float u[3] = {qb[0] - qa[0], qb[1] - qa[1], qb[2] - qa[2]}, v[3] = {qc[0] - qa[0], qc[1] - qa[1], qc[2] - qa[2]};
        const float cr[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        float num = qa[0] * cr[0];
        num += qa[1] * cr[1];
        num += qa[2] * cr[2];
        if (num == 0) return 0.0f;
        auto dot = [](const float *x, const float *y) { float r = x[0] * y[0]; r += x[1] * y[1]; r += x[2] * y[2]; return r; };
        const float den = float(1) + dot(qa, qb) + dot(qa, qc) + dot(qb, qc);
        return float(2) * atan2f_portable(num, den);
    }

    float node_solid_angle(int node, const float q0[3], float acc2) const {
        float approx[4];
        unsigned descend = 0;
        for (int i = 0; i < 4; ++i) {
            float q[3] = {q0[0] - row(node, 1)[i], q0[1] - row(node, 2)[i], q0[2] - row(node, 3)[i]};
            const float ql2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
            const bool desc = ql2 <= row(node, 0)[i] * acc2;
            const float m2 = float(1.0) / ql2, m1 = std::sqrt(m2);
            q[0] *= m1; q[1] *= m1; q[2] *= m1;
            float om = -m2 * (q[0] * row(node, 4)[i] + q[1] * row(node, 5)[i] + q[2] * row(node, 6)[i]);
            const float q2[3] = {q[0] * q[0], q[1] * q[1], q[2] * q[2]};
            const float m3 = m2 * m1;
            const float o1 = m3 * (row(node, 7)[i] + row(node, 8)[i] + row(node, 9)[i]
                                   - float(3.0) * ((q2[0] * row(node, 7)[i] + q2[1] * row(node, 8)[i] + q2[2] * row(node, 9)[i]) +
                                                   q[0] * q[1] * row(node, 10)[i] + q[0] * q[2] * row(node, 12)[i] + q[1] * q[2] * row(node, 11)[i]));
            om += o1;
            const float q3[3] = {q2[0] * q[0], q2[1] * q[1], q2[2] * q[2]};
            const float m4 = m2 * m2;
            const float t0[3] = {row(node, 20)[i] + row(node, 21)[i], row(node, 22)[i] + row(node, 17)[i], row(node, 18)[i] + row(node, 19)[i]};
            const float t1[3] = {q[1] * row(node, 17)[i] + q[2] * row(node, 18)[i], q[2] * row(node, 19)[i] + q[0] * row(node, 20)[i],
                                 q[0] * row(node, 21)[i] + q[1] * row(node, 22)[i]};
            const float da = q[0] * (float(3) * row(node, 13)[i] + t0[0]) + q[1] * (float(3) * row(node, 14)[i] + t0[1]) + q[2] * (float(3) * row(node, 15)[i] + t0[2]);
            const float db = (q3[0] * row(node, 13)[i] + q3[1] * row(node, 14)[i] + q3[2] * row(node, 15)[i]) + q[0] * q[1] * q[2] * row(node, 16)[i] +
                             (q2[0] * t1[0] + q2[1] * t1[1] + q2[2] * t1[2]);
            const float o2 = m4 * (float(1.5) * da - float(7.5) * db);
            om += o2;
            const bool use = std::isfinite(om) && !desc;
            approx[i] = use ? om : 0.0f;
            if (!use) descend |= 1u << i;
        }
        if (descend == 0xfu) {
            // (all four descend: the reference returns before evaluating the expansions)
            approx[0] = approx[1] = approx[2] = approx[3] = 0.0f;
        }
        float sum = approx[0];
        for (int i = 1; i < 4; ++i) sum += approx[i];
        float cd[4] = {0, 0, 0, 0};
        int s;
        for (s = 0; s < 4; ++s) {
            if ((descend >> s) & 1u) {
                const uint32_t c = child[4 * (size_t)node + s];
                if (c & INTERNAL) {
                    if (c == EMPTY) { descend &= (1u << s) - 1u; break; }
                    cd[s] = node_solid_angle((int)(c & ~INTERNAL), q0, acc2);
                } else {
                    cd[s] = tri_solid_angle(&U[3 * (size_t)F[3 * c]], &U[3 * (size_t)F[3 * c + 1]], &U[3 * (size_t)F[3 * c + 2]], q0);
                }
            }
        }
        float ps = (descend & 1u) ? cd[0] : 0.0f;
        for (int i = 1; i < s; ++i) ps += ((descend >> i) & 1u) ? cd[i] : 0.0f;
        return sum + ps;
    }
    int depth(int node = 0) const {
        int d = 0;
        for (int s = 0; s < 4; ++s) {
            const uint32_t c = child[4 * (size_t)node + s];
            if (c == EMPTY) break;
            if (c & INTERNAL) d = std::max(d, depth((int)(c & ~INTERNAL)));
        }
        return d + 1;
    }
    float solid_angle(const float q[3], float accuracy_scale) const {
        if (nn == 0) return 0.0f;
        return node_solid_angle(0, q, accuracy_scale * accuracy_scale);
    }
    // igl::fast_winding_number(fwn_bvh, accuracy_scale, p): float query, / (4 PI) in double
    double winding_number(const double p[3], float accuracy_scale = 2.0f) const {
        const float q[3] = {(float)p[0], (float)p[1], (float)p[2]};
        return solid_angle(q, accuracy_scale) / (4.0 * 3.1415926535897932384626433832795);
    }
};

}  // namespace host
}  // namespace svsdf
