// Generic pointwise forward(x_ref, elem_id) of the triangle model, its VJP, the deterministic fold of
// per-row contributions, the edge branch, the coords/u_full assembly and the halo pack/unpack.
// Reference: /root/reference/src/models.py:292-376 (callers src/loss.py:65,103 and src/plots.py:183-187).
#include "../../include/hidenn_b200.h"
#include "common.cuh"
#include "tri_plan.h"

namespace hidenn {

constexpr int kBlock = 256;

template <typename R> struct ElemNodes {
    typename Real2<R>::type v0, v1, v2, U0, U1, U2;
};

template <typename R>
__device__ __forceinline__ ElemNodes<R> gather_elem(const TriPlanDev& P, int64_t e, const typename Real2<R>::type* x_free,
                                                    const typename Real2<R>::type* x_fixed,
                                                    const typename Real2<R>::type* u_free,
                                                    const typename Real2<R>::type* u_fixed) {
    using R2 = typename Real2<R>::type;
    const int n0 = __ldg(P.conn32 + 3 * e), n1 = __ldg(P.conn32 + 3 * e + 1), n2 = __ldg(P.conn32 + 3 * e + 2);
    ElemNodes<R> N;
    N.v0 = load_slot<R2>(x_free, x_fixed, __ldg(P.xslot + n0));
    N.v1 = load_slot<R2>(x_free, x_fixed, __ldg(P.xslot + n1));
    N.v2 = load_slot<R2>(x_free, x_fixed, __ldg(P.xslot + n2));
    N.U0 = load_slot<R2>(u_free, u_fixed, __ldg(P.uslot + n0));
    N.U1 = load_slot<R2>(u_free, u_fixed, __ldg(P.uslot + n1));
    N.U2 = load_slot<R2>(u_free, u_fixed, __ldg(P.uslot + n2));
    return N;
}

template <typename R>
__global__ void __launch_bounds__(kBlock)
tri_eval_fwd_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                    const typename Real2<R>::type* __restrict__ x_fixed, const typename Real2<R>::type* __restrict__ u_free,
                    const typename Real2<R>::type* __restrict__ u_fixed, const typename Real2<R>::type* __restrict__ x_ref,
                    const int64_t* __restrict__ elem_id, int64_t M, typename Real2<R>::type* __restrict__ u_h,
                    R* __restrict__ detJ, R* __restrict__ grad_u) {
    using R2 = typename Real2<R>::type;
    for (int64_t m = (int64_t)blockIdx.x * kBlock + threadIdx.x; m < M; m += (int64_t)gridDim.x * kBlock) {
        const int64_t e = elem_id[m];
        const ElemNodes<R> N = gather_elem<R>(P, e, x_free, x_fixed, u_free, u_fixed);
        const R2 xr = x_ref[m];
        const R z = R(1) - xr.x - xr.y;
        if (u_h) u_h[m] = mk2<R>(xr.x * N.U0.x + xr.y * N.U1.x + z * N.U2.x, xr.x * N.U0.y + xr.y * N.U1.y + z * N.U2.y);
        // P.jinv_t (correct-math switch): J^-T instead of the reference's J^-1 = the same expressions on the transposed Jacobian
        const R a = N.v0.x - N.v2.x, d = N.v1.y - N.v2.y;
        const R b = P.jinv_t ? N.v0.y - N.v2.y : N.v1.x - N.v2.x, c = P.jinv_t ? N.v1.x - N.v2.x : N.v0.y - N.v2.y;
        const R det = a * d - b * c;
        if (detJ) detJ[m] = det;
        if (grad_u) {
            const R inv = rcp(det);
            const R j00 = d * inv, j01 = -b * inv, j10 = -c * inv, j11 = a * inv;
            const R p0 = N.U0.x - N.U2.x, p1 = N.U1.x - N.U2.x, q0 = N.U0.y - N.U2.y, q1 = N.U1.y - N.U2.y;
            R* g = grad_u + 4 * m;
            g[0] = p0 * j00 + p1 * j01; g[1] = p0 * j10 + p1 * j11;
            g[2] = q0 * j00 + q1 * j01; g[3] = q0 * j10 + q1 * j11;
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(kBlock)
tri_eval_bwd_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                    const typename Real2<R>::type* __restrict__ x_fixed, const typename Real2<R>::type* __restrict__ u_free,
                    const typename Real2<R>::type* __restrict__ u_fixed, const typename Real2<R>::type* __restrict__ x_ref,
                    const int64_t* __restrict__ elem_id, int64_t M, const typename Real2<R>::type* __restrict__ cu,
                    const R* __restrict__ cd, const R* __restrict__ cG, R* __restrict__ row_gx, R* __restrict__ row_gu) {
    using R2 = typename Real2<R>::type;
    for (int64_t m = (int64_t)blockIdx.x * kBlock + threadIdx.x; m < M; m += (int64_t)gridDim.x * kBlock) {
        const int64_t e = elem_id[m];
        const ElemNodes<R> N = gather_elem<R>(P, e, x_free, x_fixed, u_free, u_fixed);
        const R2 xr = x_ref[m];
        const R z = R(1) - xr.x - xr.y;
        const R a = N.v0.x - N.v2.x, d = N.v1.y - N.v2.y;
        const R b = P.jinv_t ? N.v0.y - N.v2.y : N.v1.x - N.v2.x, c = P.jinv_t ? N.v1.x - N.v2.x : N.v0.y - N.v2.y;
        const R det = a * d - b * c;
        const R inv = rcp(det);
        const R j00 = d * inv, j01 = -b * inv, j10 = -c * inv, j11 = a * inv;
        const R p0 = N.U0.x - N.U2.x, p1 = N.U1.x - N.U2.x, q0 = N.U0.y - N.U2.y, q1 = N.U1.y - N.U2.y;
        const R G00 = p0 * j00 + p1 * j01, G01 = p0 * j10 + p1 * j11, G10 = q0 * j00 + q1 * j01, G11 = q0 * j10 + q1 * j11;
        R c00 = R(0), c01 = R(0), c10 = R(0), c11 = R(0);
        if (cG) { c00 = cG[4 * m]; c01 = cG[4 * m + 1]; c10 = cG[4 * m + 2]; c11 = cG[4 * m + 3]; }
        const R2 ku = cu ? cu[m] : mk2<R>(R(0), R(0));
        const R kd = cd ? cd[m] : R(0);
        // M = cG . Jinv  (cotangent of dU);  dJ = kd * adj - M^T G
        const R M00 = c00 * j00 + c01 * j10, M01 = c00 * j01 + c01 * j11;
        const R M10 = c10 * j00 + c11 * j10, M11 = c10 * j01 + c11 * j11;
        const R D00 = kd * d - (M00 * G00 + M10 * G10), D11 = kd * a - (M01 * G01 + M11 * G11);
        const R Db = -kd * c - (M00 * G01 + M10 * G11), Dc = -kd * b - (M01 * G00 + M11 * G10);      // d/db, d/dc of the expression
        const R D01 = P.jinv_t ? Dc : Db, D10 = P.jinv_t ? Db : Dc;                                    // ... of the true Jacobian entries
        R* gx = row_gx + 6 * m;
        R* gu = row_gu + 6 * m;
        gx[0] = D00; gx[1] = D10; gx[2] = D01; gx[3] = D11; gx[4] = -(D00 + D01); gx[5] = -(D10 + D11);
        gu[0] = xr.x * ku.x + M00; gu[1] = xr.x * ku.y + M10;
        gu[2] = xr.y * ku.x + M01; gu[3] = xr.y * ku.y + M11;
        gu[4] = z * ku.x - (M00 + M01); gu[5] = z * ku.y - (M10 + M11);
    }
}

// rows of one element summed in the given (stable-sorted) order -> elem_tmp[e][12] = gx(6), gu(6)
template <typename R>
__global__ void __launch_bounds__(kBlock)
fold_rows_to_elems_kernel(int64_t Ne, const R* __restrict__ row_gx, const R* __restrict__ row_gu,
                          const int64_t* __restrict__ order, const int64_t* __restrict__ seg, R* __restrict__ elem_tmp) {
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e < Ne; e += (int64_t)gridDim.x * kBlock) {
        R acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = R(0);
        for (int64_t r = seg[e]; r < seg[e + 1]; ++r) {
            const int64_t m = order[r];
#pragma unroll
            for (int k = 0; k < 6; ++k) { acc[k] += row_gx[6 * m + k]; acc[6 + k] += row_gu[6 * m + k]; }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) elem_tmp[12 * e + k] = acc[k];
    }
}

// node-centric fold over the global node->element CSR (ascending element id): deterministic
template <typename R>
__global__ void __launch_bounds__(kBlock)
fold_elems_to_nodes_kernel(const TriPlanDev P, const R* __restrict__ elem_tmp, typename Real2<R>::type* __restrict__ gx_free,
                           typename Real2<R>::type* __restrict__ gu_free) {
    for (int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x; n < P.n_nodes; n += (int64_t)gridDim.x * kBlock) {
        R ax = R(0), ay = R(0), bx = R(0), by = R(0);
        for (int64_t k = P.n2e_off[n]; k < P.n2e_off[n + 1]; ++k) {
            const int ent = __ldg(P.n2e_ent + k);
            const int64_t e = ent >> 2;
            const int c = ent & 3;
            const R* t = elem_tmp + 12 * e;
            bx += t[2 * c]; by += t[2 * c + 1];
            ax += t[6 + 2 * c]; ay += t[6 + 2 * c + 1];
        }
        const int xs = __ldg(P.xslot + n), us = __ldg(P.uslot + n);
        if (gx_free && xs >= 0) gx_free[xs] = mk2<R>(bx, by);
        if (gu_free && us >= 0) gu_free[us] = mk2<R>(ax, ay);
    }
}

template <typename R>
__global__ void __launch_bounds__(kBlock)
tri_edge_fwd_kernel(const TriPlanDev P, const typename Real2<R>::type* __restrict__ x_free,
                    const typename Real2<R>::type* __restrict__ x_fixed, const typename Real2<R>::type* __restrict__ u_free,
                    const typename Real2<R>::type* __restrict__ u_fixed, const R* __restrict__ xi,
                    const int64_t* __restrict__ edge_id, int64_t M, typename Real2<R>::type* __restrict__ u_h,
                    R* __restrict__ ds) {
    using R2 = typename Real2<R>::type;
    for (int64_t m = (int64_t)blockIdx.x * kBlock + threadIdx.x; m < M; m += (int64_t)gridDim.x * kBlock) {
        const int64_t e = edge_id[m];
        const int a = __ldg(P.edges32 + 2 * e), b = __ldg(P.edges32 + 2 * e + 1);
        const R2 x0 = load_slot<R2>(x_free, x_fixed, __ldg(P.xslot + a)), x1 = load_slot<R2>(x_free, x_fixed, __ldg(P.xslot + b));
        const R2 U0 = load_slot<R2>(u_free, u_fixed, __ldg(P.uslot + a)), U1 = load_slot<R2>(u_free, u_fixed, __ldg(P.uslot + b));
        const R t = xi[m];
        u_h[m] = mk2<R>((R(1) - t) * U0.x + t * U1.x, (R(1) - t) * U0.y + t * U1.y);
        const R dx = x1.x - x0.x, dy = x1.y - x0.y;
        ds[m] = sqrt(dx * dx + dy * dy);
    }
}

template <typename R>
__global__ void __launch_bounds__(kBlock)
assemble_kernel(const int32_t* __restrict__ slot, int64_t Nn, const typename Real2<R>::type* __restrict__ free_v,
                const typename Real2<R>::type* __restrict__ fixed_v, typename Real2<R>::type* __restrict__ full) {
    using R2 = typename Real2<R>::type;
    for (int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x; n < Nn; n += (int64_t)gridDim.x * kBlock) {
        const int s = __ldg(slot + n);
        full[n] = (s >= 0 || fixed_v) ? load_slot<R2>(free_v, fixed_v, s) : mk2<R>(R(0), R(0));
    }
}

template <typename R>
__global__ void halo_pack_kernel(const typename Real2<R>::type* __restrict__ g, const int32_t* __restrict__ idx, int64_t n,
                                 typename Real2<R>::type* __restrict__ buf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] = g[idx[i]];
}
template <typename R>
__global__ void halo_unpack_kernel(typename Real2<R>::type* __restrict__ g, const int32_t* __restrict__ idx, int64_t n,
                                   const typename Real2<R>::type* __restrict__ buf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[idx[i]] = buf[i];
}

template <typename R, bool PACK>
__global__ void halo_all_kernel(typename Real2<R>::type* gx, const int32_t* __restrict__ xrows, const int32_t* __restrict__ xpos, int64_t nx,
                                typename Real2<R>::type* gu, const int32_t* __restrict__ urows, const int32_t* __restrict__ upos, int64_t nu,
                                R* loss, int64_t S, R* buf, R* zero_other) {
    using R2 = typename Real2<R>::type;
    R2* bx = reinterpret_cast<R2*>(buf) + 1;
    R2* bu = bx + S;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (PACK && zero_other) {
        R2* z = reinterpret_cast<R2*>(zero_other);
        for (int64_t i = i0; i < 1 + 2 * S; i += stride) z[i] = mk2<R>(R(0), R(0));
    }
    if (i0 == 0) {
        if (PACK) { buf[0] = *loss; buf[1] = R(0); } else *loss = buf[0];
    }
    if (gx)
        for (int64_t i = i0; i < nx; i += stride) {
            if (PACK) bx[xpos[i]] = gx[xrows[i]]; else gx[xrows[i]] = bx[xpos[i]];
        }
    if (gu)
        for (int64_t i = i0; i < nu; i += stride) {
            if (PACK) bu[upos[i]] = gu[urows[i]]; else gu[urows[i]] = bu[upos[i]];
        }
}

template <typename R, bool PACK>
static int halo_all(R* gx, const int32_t* xrows, const int32_t* xpos, int64_t nx, R* gu, const int32_t* urows, const int32_t* upos,
                    int64_t nu, R* loss, int64_t S, R* buf, R* zero_other, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(loss && buf, "halo_pack_all/unpack_all: NULL loss or buffer");
    HIDENN_REQUIRE((nx == 0 || !gx || (xrows && xpos)) && (nu == 0 || !gu || (urows && upos)), "halo_pack_all/unpack_all: NULL index arrays");
    const int64_t n = std::max<int64_t>(std::max(nx, nu), zero_other ? 1 + 2 * S : 0);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8));
    halo_all_kernel<R, PACK><<<grid, 256, 0, (cudaStream_t)s>>>((R2*)gx, xrows, xpos, nx, (R2*)gu, urows, upos, nu, loss, S, buf, zero_other);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

static inline int grid_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + kBlock - 1) / kBlock, 148 * 32)); }

template <typename R>
static int eval_fwd(const hidenn_tri_plan* p, const R* xf, const R* xb, const R* uf, const R* ub, const R* x_ref, const int64_t* eid,
                    int64_t M, R* u_h, R* detJ, R* grad_u, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p && p->device >= 0, "tri_eval_fwd: needs a device plan (no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    HIDENN_REQUIRE(M == 0 || (x_ref && eid), "tri_eval_fwd: NULL inputs");
    if (plan_ensure_generic(const_cast<hidenn_tri_plan*>(p))) return 1;
    if (M == 0) return 0;
    tri_eval_fwd_kernel<R><<<grid_for(M), kBlock, 0, (cudaStream_t)s>>>(p->dev, (const R2*)xf, (const R2*)xb, (const R2*)uf,
                                                                         (const R2*)ub, (const R2*)x_ref, eid, M, (R2*)u_h, detJ, grad_u);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int eval_bwd(const hidenn_tri_plan* p, const R* xf, const R* xb, const R* uf, const R* ub, const R* x_ref, const int64_t* eid,
                    int64_t M, const R* cu, const R* cd, const R* cG, R* row_gx, R* row_gu, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p && p->device >= 0, "tri_eval_bwd: needs a device plan (no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    HIDENN_REQUIRE(M == 0 || (x_ref && eid && row_gx && row_gu), "tri_eval_bwd: NULL inputs");
    if (plan_ensure_generic(const_cast<hidenn_tri_plan*>(p))) return 1;
    if (M == 0) return 0;
    tri_eval_bwd_kernel<R><<<grid_for(M), kBlock, 0, (cudaStream_t)s>>>(p->dev, (const R2*)xf, (const R2*)xb, (const R2*)uf,
                                                                         (const R2*)ub, (const R2*)x_ref, eid, M, (const R2*)cu, cd, cG,
                                                                         row_gx, row_gu);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int fold_rows(const hidenn_tri_plan* p, const R* row_gx, const R* row_gu, const int64_t* order, const int64_t* seg, int64_t M,
                     R* elem_tmp, R* gx, R* gu, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p && p->device >= 0, "tri_fold_rows: needs a device plan (no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    HIDENN_REQUIRE(order && seg && elem_tmp, "tri_fold_rows: NULL inputs");
    if (plan_ensure_generic(const_cast<hidenn_tri_plan*>(p))) return 1;
    (void)M;
    if (p->n_elems > 0)
        fold_rows_to_elems_kernel<R><<<grid_for(p->n_elems), kBlock, 0, (cudaStream_t)s>>>(p->n_elems, row_gx, row_gu, order, seg, elem_tmp);
    fold_elems_to_nodes_kernel<R><<<grid_for(p->n_nodes), kBlock, 0, (cudaStream_t)s>>>(p->dev, elem_tmp, (R2*)gx, (R2*)gu);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int edge_fwd(const hidenn_tri_plan* p, const R* xf, const R* xb, const R* uf, const R* ub, const R* xi, const int64_t* eid,
                    int64_t M, R* u_h, R* ds, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p && p->device >= 0, "tri_edge_fwd: needs a device plan (no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    if (plan_ensure_generic(const_cast<hidenn_tri_plan*>(p))) return 1;
    if (M == 0) return 0;
    HIDENN_REQUIRE(xi && eid && u_h && ds, "tri_edge_fwd: NULL inputs");
    tri_edge_fwd_kernel<R><<<grid_for(M), kBlock, 0, (cudaStream_t)s>>>(p->dev, (const R2*)xf, (const R2*)xb, (const R2*)uf,
                                                                         (const R2*)ub, xi, eid, M, (R2*)u_h, ds);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R>
static int assemble(const hidenn_tri_plan* p, int which, const R* free_v, const R* fixed_v, R* full, void* s) {
    using R2 = typename Real2<R>::type;
    HIDENN_REQUIRE(p && p->device >= 0, "tri_assemble: needs a device plan (no CPU fallback)");
    DeviceScope scope;
    HIDENN_CUDA_OK(scope.enter(p->device));
    HIDENN_REQUIRE(full && (which == 0 || which == 1), "tri_assemble: bad arguments");
    if (plan_ensure_generic(const_cast<hidenn_tri_plan*>(p))) return 1;
    assemble_kernel<R><<<grid_for(p->n_nodes), kBlock, 0, (cudaStream_t)s>>>(which == 0 ? p->dev.xslot : p->dev.uslot, p->n_nodes,
                                                                              (const R2*)free_v, (const R2*)fixed_v, (R2*)full);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename R> static int halo_pack(const R* g, const int32_t* idx, int64_t n, R* buf, void* s) {
    using R2 = typename Real2<R>::type;
    if (n <= 0) return 0;
    HIDENN_REQUIRE(g && idx && buf, "halo_pack: NULL");
    halo_pack_kernel<R><<<grid_for(n), kBlock, 0, (cudaStream_t)s>>>((const R2*)g, idx, n, (R2*)buf);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}
template <typename R> static int halo_unpack(R* g, const int32_t* idx, int64_t n, const R* buf, void* s) {
    using R2 = typename Real2<R>::type;
    if (n <= 0) return 0;
    HIDENN_REQUIRE(g && idx && buf, "halo_unpack: NULL");
    halo_unpack_kernel<R><<<grid_for(n), kBlock, 0, (cudaStream_t)s>>>((R2*)g, idx, n, (const R2*)buf);
    HIDENN_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace hidenn

using namespace hidenn;

#define HIDENN_EVAL_API(SUF, T)                                                                                                   \
    extern "C" int hidenn_tri_eval_fwd_##SUF(const hidenn_tri_plan* p, const T* a, const T* b, const T* c, const T* d, const T* x,  \
                                             const int64_t* e, int64_t M, T* u, T* dj, T* g, void* s) {                           \
        return eval_fwd<T>(p, a, b, c, d, x, e, M, u, dj, g, s);                                                                  \
    }                                                                                                                             \
    extern "C" int hidenn_tri_eval_bwd_##SUF(const hidenn_tri_plan* p, const T* a, const T* b, const T* c, const T* d, const T* x,  \
                                             const int64_t* e, int64_t M, const T* cu, const T* cd, const T* cG, T* rx, T* ru,    \
                                             void* s) {                                                                           \
        return eval_bwd<T>(p, a, b, c, d, x, e, M, cu, cd, cG, rx, ru, s);                                                        \
    }                                                                                                                             \
    extern "C" int hidenn_tri_fold_rows_##SUF(const hidenn_tri_plan* p, const T* rx, const T* ru, const int64_t* o,               \
                                              const int64_t* sg, int64_t M, T* tmp, T* gx, T* gu, void* s) {                      \
        return fold_rows<T>(p, rx, ru, o, sg, M, tmp, gx, gu, s);                                                                 \
    }                                                                                                                             \
    extern "C" int hidenn_tri_edge_fwd_##SUF(const hidenn_tri_plan* p, const T* a, const T* b, const T* c, const T* d, const T* xi, \
                                             const int64_t* e, int64_t M, T* u, T* ds, void* s) {                                 \
        return edge_fwd<T>(p, a, b, c, d, xi, e, M, u, ds, s);                                                                    \
    }                                                                                                                             \
    extern "C" int hidenn_tri_assemble_##SUF(const hidenn_tri_plan* p, int which, const T* fr, const T* fx, T* full, void* s) {   \
        return assemble<T>(p, which, fr, fx, full, s);                                                                            \
    }                                                                                                                             \
    extern "C" int hidenn_halo_pack_##SUF(const T* g, const int32_t* idx, int64_t n, T* buf, void* s) {                           \
        return halo_pack<T>(g, idx, n, buf, s);                                                                                   \
    }                                                                                                                             \
    extern "C" int hidenn_halo_unpack_##SUF(T* g, const int32_t* idx, int64_t n, const T* buf, void* s) {                         \
        return halo_unpack<T>(g, idx, n, buf, s);                                                                                 \
    }

#define HIDENN_HALO_ALL_API(SUF, T)                                                                                              \
    extern "C" int hidenn_halo_pack_all_##SUF(const T* gx, const int32_t* xr, const int32_t* xp, int64_t nx, const T* gu,          \
                                              const int32_t* ur, const int32_t* up, int64_t nu, const T* loss, int64_t S, T* buf,  \
                                              T* zero_other, void* s) {                                                           \
        return halo_all<T, true>(const_cast<T*>(gx), xr, xp, nx, const_cast<T*>(gu), ur, up, nu, const_cast<T*>(loss), S, buf,     \
                                 zero_other, s);                                                                                  \
    }                                                                                                                             \
    extern "C" int hidenn_halo_unpack_all_##SUF(T* gx, const int32_t* xr, const int32_t* xp, int64_t nx, T* gu, const int32_t* ur, \
                                                const int32_t* up, int64_t nu, T* loss, int64_t S, const T* buf, void* s) {        \
        return halo_all<T, false>(gx, xr, xp, nx, gu, ur, up, nu, loss, S, const_cast<T*>(buf), nullptr, s);                                \
    }
HIDENN_HALO_ALL_API(f64, double)
HIDENN_HALO_ALL_API(f32, float)

HIDENN_EVAL_API(f64, double)
HIDENN_EVAL_API(f32, float)
