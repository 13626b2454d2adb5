// Device code shared by the tile kernels: per-element closed form, kernel constants, shared-memory record layouts.
// Per-element formulas: SURVEY.md Appendix A.1 with  d psi/dJ = -(P Jinv)^T G  (/root/reference/src/models.py:316-357,
// /root/reference/src/loss.py:55-88).
#pragma once
#include "../../include/hidenn_b200.h"
#include "common.cuh"

namespace hidenn {

constexpr int kTileBlock = 256;

template <typename R> struct TriConsts {
    R c00, c01, c02, c11, c12, c22, W;
    R fb[6];
};

// 1/x to full precision without the slow-path branch of the IEEE division: hardware seed + 2 Newton steps.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }

// BODY: a body force table is present (Fb != 0).  ISO: C has the plane-stress form c02 = c12 = 0.
//
// Division-free chain: with adj = [[d,-b],[-c,a]] (J^-1 = adj/det) every quantity is a numerator times a power of
// 1/det:  G = Gt/det,  eps = et/det,  sigma = st/det,  psi = pt/det^2,  |det| W psi = k1 pt  with k1 = W/|det|,
// |det| W P J^-1 = k1 Mt,  and  dE/dJ = k2 (pt adj' - Mt^T Gt)  with k2 = k1/det.  The reciprocal (MUFU seed + 2
// Newton steps, 7 dependent operations) is needed only for the three final scalings, so it overlaps the ~40
// numerator operations instead of heading the dependency chain; the FP64 count drops from ~85 to ~68 per element.
template <typename R, bool BODY, bool ISO>
__device__ __forceinline__ void tri_element(const typename Real2<R>::type v0, const typename Real2<R>::type v1,
                                            const typename Real2<R>::type v2, const typename Real2<R>::type U0,
                                            const typename Real2<R>::type U1, const typename Real2<R>::type U2,
                                            const TriConsts<R>& K, R& energy, typename Real2<R>::type gu[3],
                                            typename Real2<R>::type gx[3], const bool jt = false) {
    // jt (correct-math switch, default off): the energy with J^-T is the reference's J^-1 expression evaluated on the
    // transposed Jacobian (b and c exchanged; det is the same), so its derivatives w.r.t. b and c come out exchanged
    const R a = v0.x - v2.x, d = v1.y - v2.y;
    const R b_in = v1.x - v2.x, c_in = v0.y - v2.y;
    const R b = jt ? c_in : b_in, c = jt ? b_in : c_in;
    const R det = a * d - b * c;
    const R inv = fast_rcp(det);
    const R p0 = U0.x - U2.x, p1 = U1.x - U2.x, q0 = U0.y - U2.y, q1 = U1.y - U2.y;
    // Gt = dU . adj^T   (the reference's J^-1 quirk: G = dU . J^-T)
    const R G00 = p0 * d - p1 * b, G01 = p1 * a - p0 * c;
    const R G10 = q0 * d - q1 * b, G11 = q1 * a - q0 * c;
    const R e0 = G00, e1 = G11, e2 = G01 + G10;
    R s0, s1, s2;
    if (ISO) {
        s0 = K.c00 * e0 + K.c01 * e1;
        s1 = K.c01 * e0 + K.c11 * e1;
        s2 = K.c22 * e2;
    } else {
        s0 = K.c00 * e0 + K.c01 * e1 + K.c02 * e2;
        s1 = K.c01 * e0 + K.c11 * e1 + K.c12 * e2;
        s2 = K.c02 * e0 + K.c12 * e1 + K.c22 * e2;
    }
    const R pt = R(0.5) * (e0 * s0 + e1 * s1 + e2 * s2);
    // Mt = St . adj   (St = [[s0,s2],[s2,s1]])
    const R M00 = s0 * d - s2 * c, M01 = s2 * a - s0 * b;
    const R M10 = s2 * d - s1 * c, M11 = s1 * a - s2 * b;
    // Nt = pt adj' - Mt^T Gt
    const R N00 = pt * d - (M00 * G00 + M10 * G10), N01 = -(pt * c) - (M00 * G01 + M10 * G11);
    const R N10 = -(pt * b) - (M01 * G00 + M11 * G10), N11 = pt * a - (M01 * G01 + M11 * G11);
    const R k1 = K.W * fabs(inv);
    const R k2 = k1 * inv;
    energy = k1 * pt;
    R g00 = k1 * M00, g10 = k1 * M10, g01 = k1 * M01, g11 = k1 * M11;
    R D00 = k2 * N00, D01 = k2 * N01, D10 = k2 * N10, D11 = k2 * N11;
    if (BODY) {
        const R A = fabs(det);
        const R bw = U0.x * K.fb[0] + U0.y * K.fb[1] + U1.x * K.fb[2] + U1.y * K.fb[3] + U2.x * K.fb[4] + U2.y * K.fb[5];
        energy -= A * bw;
        const R sb = det < R(0) ? bw : -bw;            // d(-|det| bw)/dJ = -sign(det) bw adj'
        D00 += sb * d; D01 -= sb * c; D10 -= sb * b; D11 += sb * a;
        gu[0] = mk2<R>(g00 - A * K.fb[0], g10 - A * K.fb[1]);
        gu[1] = mk2<R>(g01 - A * K.fb[2], g11 - A * K.fb[3]);
        gu[2] = mk2<R>(-(g00 + g01) - A * K.fb[4], -(g10 + g11) - A * K.fb[5]);
    } else {
        gu[0] = mk2<R>(g00, g10);
        gu[1] = mk2<R>(g01, g11);
        gu[2] = mk2<R>(-(g00 + g01), -(g10 + g11));
    }
    {       // D01 = d/db, D10 = d/dc of the expression above: with jt they are d/dc and d/db of the true Jacobian entries
        const R t01 = jt ? D10 : D01, t10 = jt ? D01 : D10;
        D01 = t01; D10 = t10;
    }
    gx[0] = mk2<R>(D00, D10);
    gx[1] = mk2<R>(D01, D11);
    gx[2] = mk2<R>(-(D00 + D01), -(D10 + D11));
}

template <typename R, bool BODY> __device__ __forceinline__ TriConsts<R> load_consts(const R* __restrict__ consts) {
    TriConsts<R> K;
    K.c00 = __ldg(consts + HIDENN_TRI_C00); K.c01 = __ldg(consts + HIDENN_TRI_C01); K.c02 = __ldg(consts + HIDENN_TRI_C02);
    K.c11 = __ldg(consts + HIDENN_TRI_C11); K.c12 = __ldg(consts + HIDENN_TRI_C12); K.c22 = __ldg(consts + HIDENN_TRI_C22);
    K.W = __ldg(consts + HIDENN_TRI_W);
#pragma unroll
    for (int k = 0; k < 6; ++k) K.fb[k] = BODY ? __ldg(consts + HIDENN_TRI_FB + k) : R(0);
    return K;
}

__device__ __forceinline__ void cp_async_pair(void* smem_dst, const void* gsrc, const int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    // 16-byte copies bypass L1 (.cg): a tile's rows are used once per CTA, L1 hit rate was 7 % (-1 % kernel time)
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory layouts of the persistent kernel.  FP64: node pairs xy | uv and partial pairs gu | gx as separate
// 16-byte arrays.  FP32: one 16-byte record (x,y,ux,uy) per node and (gu.x,gu.y,gx.x,gx.y) per fold slot, so every
// access is a single 128-bit pass of 8 lanes in both precisions.
template <typename R> struct NodeBuf;
template <> struct NodeBuf<double> {
    double2* xy; double2* uv;
    __device__ __forceinline__ NodeBuf(void* base, int max_local) : xy((double2*)base), uv((double2*)base + max_local) {}
    __device__ __forceinline__ void load(unsigned l, double2& a, double2& b) const { a = xy[l]; b = uv[l]; }
    __device__ __forceinline__ void* xy_ptr(int i) const { return xy + i; }
    __device__ __forceinline__ void* uv_ptr(int i) const { return uv + i; }
};
template <> struct NodeBuf<float> {
    float4* rec;
    __device__ __forceinline__ NodeBuf(void* base, int) : rec((float4*)base) {}
    __device__ __forceinline__ void load(unsigned l, float2& a, float2& b) const {
        const float4 r = rec[l];
        a = make_float2(r.x, r.y); b = make_float2(r.z, r.w);
    }
    __device__ __forceinline__ void* xy_ptr(int i) const { return reinterpret_cast<float2*>(rec + i); }
    __device__ __forceinline__ void* uv_ptr(int i) const { return reinterpret_cast<float2*>(rec + i) + 1; }
};
template <typename R> struct PartBuf;
template <> struct PartBuf<double> {
    double2* pu; double2* px;
    __device__ __forceinline__ PartBuf(void* base, int n) : pu((double2*)base), px((double2*)base + n) {}
    __device__ __forceinline__ void store(unsigned p, double2 gu, double2 gx) const { pu[p] = gu; px[p] = gx; }
    __device__ __forceinline__ void load(unsigned p, double2& gu, double2& gx) const { gu = pu[p]; gx = px[p]; }
};
template <> struct PartBuf<float> {
    float4* rec;
    __device__ __forceinline__ PartBuf(void* base, int) : rec((float4*)base) {}
    __device__ __forceinline__ void store(unsigned p, float2 gu, float2 gx) const { rec[p] = make_float4(gu.x, gu.y, gx.x, gx.y); }
    __device__ __forceinline__ void load(unsigned p, float2& gu, float2& gx) const {
        const float4 r = rec[p];
        gu = make_float2(r.x, r.y); gx = make_float2(r.z, r.w);
    }
};

}  // namespace hidenn
