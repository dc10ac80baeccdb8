// oracle/ref_fwn_shim.cpp — TEST INFRASTRUCTURE ONLY.
// Thin C entry points over the REFERENCE's own fast-winding-number code, compiled from the source where it lies
// (/root/reference/src/utils/include/igl/FastWindingNumberForSoups.h — the HDK UT_SolidAngle<float,float> that
// igl::fast_winding_number wraps; it needs no Eigen).  Built by `make ref` into oracle/_ref/libref_fwn.so (git-ignored).
// The two functions repeat what the wrappers do around it:
//   igl::fast_winding_number(V, F, order, fwn_bvh)        fast_winding_number.cpp:380-408  (double -> float copies, init)
//   igl::fast_winding_number(fwn_bvh, accuracy_scale, p)  fast_winding_number.cpp:439-457  (float query, / (4*PI) in double)
// Used only by tests/ (and tests/golden/make_fwn_golden.py) to pin the oracle's exact winding number against the
// reference's float approximation of it.  No reference source is copied into this repository.
#include <cstdint>
#include <cstring>
#include <vector>

#include "igl/FastWindingNumberForSoups.h"

namespace {
using namespace igl::FastWindingNumber::HDK_Sample;
struct RefFwn {
    UT_SolidAngle<float, float> ut_solid_angle;
    std::vector<UT_Vector3T<float>> U;
    std::vector<int> F;
};
}  // namespace

extern "C" {

void *ref_fwn_create(const double *V, int nv, const int *F, int nf, int order) {
    RefFwn *r = new RefFwn();
    r->U.resize(nv);
    for (int i = 0; i < nv; ++i)
        for (int j = 0; j < 3; ++j) r->U[i][j] = (float)V[3 * i + j];
    r->F.assign(F, F + 3 * (size_t)nf);
    r->ut_solid_angle.clear();
    r->ut_solid_angle.init(nf, &r->F[0], nv, &r->U[0], order);
    return r;
}

void ref_fwn_destroy(void *h) { delete (RefFwn *)h; }

// Structure dump (compiled with -fno-access-control): the 4-way BVH the reference built and its per-node expansion data, so
// that the product's own builder (csrc/host/fwn_bvh.hpp) can be compared node by node, not only through winding numbers.
// children: [nn][4] raw child words (EMPTY = 0xffffffff, bit 31 = internal); boxdata: [nn][23][4] floats in BoxData member order
// (myMaxPDist2, myAverageP[3], myN[3], myNijDiag[3], myNxy_Nyx, myNyz_Nzy, myNzx_Nxz, myNijkDiag[3], mySumPermuteNxyz,
//  my2Nxxy_Nyxx, my2Nxxz_Nzxx, my2Nyyz_Nzyy, my2Nyyx_Nxyy, my2Nzzx_Nxzz, my2Nzzy_Nyzz), 4 lanes = the node's 4 children.
int ref_fwn_num_nodes(void *h) { return (int)((RefFwn *)h)->ut_solid_angle.myTree.getNumNodes(); }
void ref_fwn_dump(void *h, uint32_t *children, float *boxdata) {
    RefFwn *r = (RefFwn *)h;
    const int nn = (int)r->ut_solid_angle.myTree.getNumNodes();
    const auto *nodes = r->ut_solid_angle.myTree.getNodes();
    for (int i = 0; i < nn; ++i)
        for (int c = 0; c < 4; ++c) children[4 * i + c] = nodes[i].child[c];
    static_assert(sizeof(r->ut_solid_angle.myData[0]) == 23 * 16, "BoxData layout");
    std::memcpy(boxdata, r->ut_solid_angle.myData.get(), (size_t)nn * 23 * 16);
}

void ref_fwn_eval(void *h, float accuracy_scale, int64_t n, const double *q, double *w_out) {
    RefFwn *r = (RefFwn *)h;
    const double PI = 3.1415926535897932384626433832795;  // igl::PI (igl/PI.h)
    for (int64_t i = 0; i < n; ++i) {
        UT_Vector3T<float> Qp;
        Qp[0] = (float)q[3 * i];
        Qp[1] = (float)q[3 * i + 1];
        Qp[2] = (float)q[3 * i + 2];
        w_out[i] = r->ut_solid_angle.computeSolidAngle(Qp, accuracy_scale) / (4.0 * PI);
    }
}
}
