// sb2_interp.cuh -- region_model::interpolate on the device: IDW for all five variables and Bayesian
// temperature kriging (BTK), producing forcing windows laid out [time][cell].
//
// Follows:
//   inverse_distance::run_interpolation     core/inverse_distance.h:142-250 (neighbour lists :160-203, per-step mean :214-249)
//   temperature_gradient_scale_computer     core/inverse_distance.h:265-325
//   model transforms                         core/inverse_distance.h:365-472
//   geo_point::distance_measure              core/geo_point.h:41-43
//   btk_interpolation                        core/bayesian_kriging.h:280-402 (covariances :93-125)
#pragma once
#include <stdint.h>

#include "sb2_math.cuh"

namespace sb2 {

enum { IDW_TEMPERATURE = 0, IDW_PRECIPITATION = 1, IDW_RADIATION = 2, IDW_WIND_SPEED = 3, IDW_REL_HUM = 4 };

struct IdwParam {
    int max_members;
    double max_distance, distance_measure_factor, zscale, default_temp_gradient, scale_factor;
    int gradient_by_equation;
};

__device__ __forceinline__ double distance_measure(double ax, double ay, double az, double bx, double by, double bz, double p, double zscale) {
    // explicit round-to-nearest products and sums: never contracted to FMA, so weights (and with them the neighbour
    // ordering) are bit-identical to the host arithmetic of the reference
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(__dmul_rn(__dmul_rn(dz, dz), zscale), zscale));
    const double e = p / 2.0;
    return sb_pow(d2, e);  // exact for e == 1 (the default distance_measure_factor 2)
}

// Step 1 (:160-203): per destination cell the sources with weight >= min_weight; if more than max_members, the
// max_members heaviest in descending weight order (std::partial_sort; ties, which the reference leaves
// implementation-defined, resolve to the lower source index here), otherwise source order.
// Output: nb_idx/nb_w/nb_f [k][cell], nb_n [cell]; nb_f = per-neighbour multiplicative transform factor.
__global__ void idw_build_neighbours_kernel(int kind, int64_t n_cells, const double* __restrict__ cx, const double* __restrict__ cy,
                                            const double* __restrict__ cz, const double* __restrict__ cslope, int n_src,
                                            const double* __restrict__ sxyz, IdwParam p, double min_weight, int32_t* __restrict__ nb_idx,
                                            double* __restrict__ nb_w, double* __restrict__ nb_f, int32_t* __restrict__ nb_n) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double x = cx[c], y = cy[c], z = cz[c];
    auto weight = [&](int k) {
        const double w = dmin(1.0, 1.0 / distance_measure(x, y, z, sxyz[3 * k], sxyz[3 * k + 1], sxyz[3 * k + 2], p.distance_measure_factor, p.zscale));
        return w;
    };
    int n_ok = 0;
    for (int k = 0; k < n_src; ++k)
        if (weight(k) >= min_weight) ++n_ok;
    int cnt = 0;
    if (n_ok <= p.max_members) {
        for (int k = 0; k < n_src; ++k) {
            const double w = weight(k);
            if (w >= min_weight) { nb_idx[(int64_t)cnt * n_cells + c] = k; nb_w[(int64_t)cnt * n_cells + c] = w; ++cnt; }
        }
    } else {
        double last_w = 2.0;  // weights are <= 1
        int last_k = -1;
        for (cnt = 0; cnt < p.max_members; ++cnt) {
            double best_w = -1.0;
            int best_k = -1;
            for (int k = 0; k < n_src; ++k) {
                const double w = weight(k);
                if (!(w >= min_weight)) continue;
                const bool after_last = (w < last_w) || (w == last_w && k > last_k);
                if (after_last && w > best_w) { best_w = w; best_k = k; }
            }
            nb_idx[(int64_t)cnt * n_cells + c] = best_k;
            nb_w[(int64_t)cnt * n_cells + c] = best_w;
            last_w = best_w;
            last_k = best_k;
        }
    }
    nb_n[c] = cnt;
    for (int j = 0; j < cnt; ++j) {
        const int k = nb_idx[(int64_t)j * n_cells + c];
        double f = 1.0;
        if (kind == IDW_PRECIPITATION) f = sb_pow(p.scale_factor, (z - sxyz[3 * k + 2]) / 100.0);  // :422-426
        else if (kind == IDW_RADIATION) f = cslope[c];                                            // :392-394
        else if (kind == IDW_TEMPERATURE) f = sxyz[3 * k + 2];                                     // station height, read by the gradient (:285-316) and the transform (:367-369)
        nb_f[(int64_t)j * n_cells + c] = f;
    }
}

// 3x3 solve with partial pivoting (arma::solve at inverse_distance.h:296-300); false when singular
__device__ inline bool solve3(double A[3][3], double b[3], double x[3]) {
    int piv[3] = {0, 1, 2};
    double scale = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = dmax(scale, fabs(A[i][j]));
    if (!(scale > 0)) return false;
    for (int c = 0; c < 3; ++c) {
        int best = c;
        for (int r = c + 1; r < 3; ++r) if (fabs(A[piv[r]][c]) > fabs(A[piv[best]][c])) best = r;
        int tmp = piv[c]; piv[c] = piv[best]; piv[best] = tmp;
        const double d = A[piv[c]][c];
        if (fabs(d) < 1e-14 * scale) return false;
        for (int r = c + 1; r < 3; ++r) {
            const double f = A[piv[r]][c] / d;
            for (int j = c; j < 3; ++j) A[piv[r]][j] -= f * A[piv[c]][j];
            b[piv[r]] -= f * b[piv[c]];
        }
    }
    for (int c = 2; c >= 0; --c) {
        double s = b[piv[c]];
        for (int j = c + 1; j < 3; ++j) s -= A[piv[c]][j] * x[j];
        x[c] = s / A[piv[c]][c];
    }
    return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
}

// Step 2 (:214-249): out[(i)*n_cells + c] = sum_k w*transform(v_k) / sum_k w over the finite neighbours, in list order.
// One thread per cell.  The block stages (a) its cells' neighbour lists -- index, weight and the per-neighbour constant
// (precipitation/radiation factor, or the station height for temperature) -- once, laid out [j][thread] so the per-step
// walk is conflict-free, and (b) the station values of a tile of steps; every step then runs out of shared memory.
// dynamic smem: tile_steps*n_src doubles + max_k*IDW_BLOCK*(2 doubles + 1 int)
constexpr int IDW_BLOCK = 128;
template <int KIND>
__global__ void __launch_bounds__(IDW_BLOCK) idw_apply_kernel(int64_t n_cells, const double* __restrict__ cz, int n_src,
                                                              const double* __restrict__ sxyz, const double* __restrict__ src /* [T][n_src] */,
                                                              int64_t first_step, int n_steps, IdwParam p, const int32_t* __restrict__ nb_idx,
                                                              const double* __restrict__ nb_w, const double* __restrict__ nb_f,
                                                              const int32_t* __restrict__ nb_n, const uint8_t* __restrict__ active,
                                                              double* __restrict__ out, int tile_steps, int max_k) {
    extern __shared__ double smem[];
    double* sv = smem;                                   // [tile_steps][n_src]
    double* sw = sv + (size_t)tile_steps * n_src;        // [max_k][IDW_BLOCK]
    double* sf = sw + (size_t)max_k * IDW_BLOCK;         // [max_k][IDW_BLOCK]
    int* sk = (int*)(sf + (size_t)max_k * IDW_BLOCK);    // [max_k][IDW_BLOCK]
    const int tid = threadIdx.x;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + tid;
    const bool ok = c < n_cells && (active == nullptr || active[c] != 0);  // cells outside the calculation filter keep their NaN fill (region_model.h:420-423)
    const int cnt = ok ? nb_n[c] : 0;
    const double z = ok ? cz[c] : 0.0;
    for (int j = 0; j < cnt; ++j) {
        sk[j * IDW_BLOCK + tid] = nb_idx[(int64_t)j * n_cells + c];
        sw[j * IDW_BLOCK + tid] = nb_w[(int64_t)j * n_cells + c];
        sf[j * IDW_BLOCK + tid] = nb_f[(int64_t)j * n_cells + c];
    }
    for (int t0 = 0; t0 < n_steps; t0 += tile_steps) {
        const int nt = min(tile_steps, n_steps - t0);
        __syncthreads();
        for (int e = tid; e < nt * n_src; e += blockDim.x) sv[e] = src[(first_step + t0) * n_src + e];
        __syncthreads();
        if (!ok) continue;
        for (int i = 0; i < nt; ++i) {
            const double* v = sv + i * n_src;
            double scale = 1.0;
            if (KIND == IDW_TEMPERATURE) {  // temperature_gradient_scale_computer::compute over the finite neighbours (:285-316)
                scale = p.default_temp_gradient;
                int nv = 0, mn = -1, mx = -1, first[4] = {-1, -1, -1, -1};
                double zmn = 0, zmx = 0;
                for (int j = 0; j < cnt; ++j) {
                    const int k = sk[j * IDW_BLOCK + tid];
                    if (!isfinite(v[k])) continue;
                    const double h = sf[j * IDW_BLOCK + tid];
                    if (nv < 4) first[nv] = k;
                    if (nv == 0) { mn = mx = k; zmn = zmx = h; }
                    else if (h < zmn) { mn = k; zmn = h; }
                    else if (h > zmx) { mx = k; zmx = h; }
                    ++nv;
                }
                bool solved = false;
                if (p.gradient_by_equation && nv > 3) {
                    double A[3][3], b[3], x[3];
                    for (int r = 0; r < 3; ++r) {
                        A[r][0] = sxyz[3 * first[r + 1]] - sxyz[3 * first[0]];
                        A[r][1] = sxyz[3 * first[r + 1] + 1] - sxyz[3 * first[0] + 1];
                        A[r][2] = sxyz[3 * first[r + 1] + 2] - sxyz[3 * first[0] + 2];
                        b[r] = v[first[r + 1]] - v[first[0]];
                    }
                    if (solve3(A, b, x)) { scale = x[2]; solved = true; }
                }
                if (!solved && nv > 1) {
                    const double dz = zmx - zmn;
                    scale = dz > 50.0 ? (v[mx] - v[mn]) / dz : p.default_temp_gradient;
                }
            }
            double sum_w = 0, sum_wv = 0;
            for (int j = 0; j < cnt; ++j) {
                const double s = v[sk[j * IDW_BLOCK + tid]];
                if (isfinite(s)) {
                    const double w = sw[j * IDW_BLOCK + tid];
                    double tv;
                    if (KIND == IDW_TEMPERATURE) tv = s + scale * (z - sf[j * IDW_BLOCK + tid]);
                    else if (KIND == IDW_PRECIPITATION || KIND == IDW_RADIATION) tv = s * sf[j * IDW_BLOCK + tid];
                    else tv = s;
                    sum_wv += w * tv;
                    sum_w += w;
                }
            }
            out[(int64_t)(t0 + i) * n_cells + c] = sum_wv / sum_w;
        }
    }
}

// ---- BTK ---------------------------------------------------------------------------------------
// k[s][c] = (sill - nug) * exp(-zscaled_distance(s, c) / range)   (bayesian_kriging.h:41-56,112-124)
// omega[c][s] = sum_j k[j][c] * K_inv[j][s]                        (:316, omega = k.t()*K_inv)
// bm[c][0..1] = ((f - F.t()*K_inv*k).t() * (I - GH_inv))[c]        (:314)
// FtKinv [2][S] = F.t()*K_inv and M22 = (I - GH_inv) come from the host (S x S algebra).
__global__ void btk_build_operators_kernel(int64_t n_cells, const double* __restrict__ cx, const double* __restrict__ cy,
                                           const double* __restrict__ cz, int n_src, const double* __restrict__ sxyz,
                                           const double* __restrict__ K_inv, const double* __restrict__ FtKinv,
                                           const double* __restrict__ M22, double sill_m_nug, double range, double zscale,
                                           double* __restrict__ omega /* [S][cells] */, double* __restrict__ bm /* [2][cells] */,
                                           double* __restrict__ kbuf /* [S][cells] scratch */) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double x = cx[c], y = cy[c], z = cz[c];
    for (int s = 0; s < n_src; ++s) {
        const double dx = sxyz[3 * s] - x, dy = sxyz[3 * s + 1] - y, dz = sxyz[3 * s + 2] - z;
        const double d = sqrt(dx * dx + dy * dy + dz * dz * zscale * zscale);
        kbuf[(int64_t)s * n_cells + c] = sill_m_nug * sb_exp(-d / range);
    }
    double g0 = 0.0, g1 = 0.0;  // (F.t()*K_inv*k)[:, c]
    for (int j = 0; j < n_src; ++j) {
        const double kj = kbuf[(int64_t)j * n_cells + c];
        g0 += FtKinv[j] * kj;
        g1 += FtKinv[n_src + j] * kj;
    }
    const double d0 = 1.0 - g0, d1 = z - g1;  // f - ...
    bm[c] = d0 * M22[0] + d1 * M22[2];
    bm[n_cells + c] = d0 * M22[1] + d1 * M22[3];
    for (int s = 0; s < n_src; ++s) {
        double acc = 0.0;
        for (int j = 0; j < n_src; ++j) acc += kbuf[(int64_t)j * n_cells + c] * K_inv[j * n_src + s];
        omega[(int64_t)s * n_cells + c] = acc;
    }
}

// per-step quantities (:384-394): beta_hat = E_beta_w * T_obs, resid = T_obs - F*beta_hat; one thread per step
__global__ void btk_step_prepare_kernel(int n_steps, int64_t first_step, int n_src, const double* __restrict__ src,
                                        const int32_t* __restrict__ valid_idx, int n_valid, const double* __restrict__ E_beta_w /* [2][n_valid] */,
                                        const double* __restrict__ sz_valid, double* __restrict__ beta /* [n_steps][2] */,
                                        double* __restrict__ resid /* [n_steps][n_valid] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    const double* v = src + (first_step + i) * n_src;
    double b0 = 0.0, b1 = 0.0;
    for (int s = 0; s < n_valid; ++s) {
        const double t = v[valid_idx[s]];
        b0 += E_beta_w[s] * t;
        b1 += E_beta_w[n_valid + s] * t;
    }
    beta[2 * i] = b0;
    beta[2 * i + 1] = b1;
    for (int s = 0; s < n_valid; ++s) resid[(int64_t)i * n_valid + s] = v[valid_idx[s]] - (1.0 * b0 + sz_valid[s] * b1);
}

// T_hat[c] = f_c.t()*beta_hat + omega[c,:]*resid - BM[c,:]*(beta_hat - E_beta_pri)   (:394-396); E_beta_pri = (0, prior_gradient(t))
__global__ void __launch_bounds__(128) btk_apply_kernel(int64_t n_cells, const double* __restrict__ cz, int n_valid,
                                                        const double* __restrict__ omega /* [n_valid][cells] */, const double* __restrict__ bm,
                                                        const double* __restrict__ beta, const double* __restrict__ resid,
                                                        const double* __restrict__ prior_gradient /* [n_steps] */, int n_steps,
                                                        const uint8_t* __restrict__ active, double* __restrict__ out /* [n_steps][cells] */,
                                                        int tile_steps) {
    extern __shared__ double sr[];  // [tile_steps][n_valid]
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = c < n_cells && (active == nullptr || active[c] != 0);
    const double z = ok ? cz[c] : 0.0, bm0 = ok ? bm[c] : 0.0, bm1 = ok ? bm[n_cells + c] : 0.0;
    for (int t0 = 0; t0 < n_steps; t0 += tile_steps) {
        const int nt = min(tile_steps, n_steps - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * n_valid; e += blockDim.x) sr[e] = resid[(int64_t)t0 * n_valid + e];
        __syncthreads();
        if (!ok) continue;
        for (int i = 0; i < nt; ++i) {
            const double b0 = beta[2 * (t0 + i)], b1 = beta[2 * (t0 + i) + 1];
            double acc = 0.0;
            for (int s = 0; s < n_valid; ++s) acc += omega[(int64_t)s * n_cells + c] * sr[i * n_valid + s];
            const double t_hat = (1.0 * b0 + z * b1) + acc;
            out[(int64_t)(t0 + i) * n_cells + c] = t_hat - (bm0 * (b0 - 0.0) + bm1 * (b1 - prior_gradient[t0 + i]));
        }
    }
}

// ---- BTK on the FP64 tensor cores ---------------------------------------------------------------------------
// The per-step kriging term  sum_s omega[c][s] * resid[t][s]  is a dense (time x stations) x (stations x cells) contraction.
// D[t][c] tiles of 8 steps x 8 cells are accumulated with mma.sync.m8n8k4.f64 (SASS DMMA): A = resid (row-major, staged per
// block in shared memory, row stride padded to 4 mod 16 doubles so a half-warp's 4 rows x 4 columns hit 16 distinct 8-byte
// banks), B = omega^T held in registers for the whole launch (it does not change over time), then the affine terms of
// bayesian_kriging.h:394-396 are applied in the epilogue.  Each warp owns 8*NT consecutive cells; a quad of lanes stores 8
// consecutive cells (64 B) per step.  KSTEPS*4 >= n_valid stations (zero padded).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// MODE 1: Bayesian temperature kriging (A = resid, epilogue of bayesian_kriging.h:394-396).
// MODE 0: inverse distance weighting with every station value finite: out[t][c] = sum_s W[c][s] * v[t][s] + addc[c], where the
//         dense operator W (idw_build_dense_kernel) folds weights, normalisation, the per-neighbour factor and, for
//         temperature, the min/max-height gradient (inverse_distance.h:304-315) into one matrix -- the "(cells x stations) x
//         (stations x steps)" contraction of the interpolation step.  A = the station series themselves.
// Staging: the rows of A for a tile of steps are contiguous in global memory (n_valid doubles each); warp 0 issues one TMA bulk copy
// (cp.async.bulk, global -> shared, completion counted on an mbarrier) per row into the padded tile, two tiles in flight (double
// buffer), so the copy of tile i+1 overlaps the DMMAs of tile i.  (Before: a load -> store loop between two __syncthreads took
// half of the kernel's stall samples, profiles/README.md.)  use_tma = 0 keeps the plain loop for rows that are not 16-byte multiples.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1 << 24)) __trap();  // a copy that never lands must not hang the device
    }
}

// KSTEPS = k-steps (groups of four k-slots) the operator fragments cover; KROW = groups of four stations a staged row holds (the row
// always carries every valid station; with COMPACT the k-slots index into it through the union list, so KSTEPS may be far smaller
// than KROW -- fewer registers, more resident warps); TILE_STEPS = steps staged per buffer.
__host__ __device__ constexpr int dense_tile_steps(int krow, int tile_steps) { return tile_steps > 0 ? tile_steps : (krow > 16 ? 32 : 64); }
__host__ __device__ constexpr int dense_smem_bytes(int krow, int tile_steps) { return 2 * dense_tile_steps(krow, tile_steps) * (krow * 4 + 4) * 8; }
template <int KSTEPS, int NT, int MODE, bool COMPACT, int KROW = KSTEPS, int TILE_STEPS = 0>
__global__ void __launch_bounds__(128) dense_apply_dmma_kernel(int64_t n_cells, const double* __restrict__ cz, int n_valid,
                                                               const double* __restrict__ omega /* [n_valid][cells] */, const double* __restrict__ bm,
                                                               const double* __restrict__ beta /* [n_steps][2] */,
                                                               const double* __restrict__ resid /* [n_steps][row_stride] */, int64_t row_stride,
                                                               const double* __restrict__ prior_gradient /* [n_steps] */, int n_steps,
                                                               const uint8_t* __restrict__ active, double* __restrict__ out /* [n_steps][cells] */,
                                                               const int32_t* __restrict__ ulist, const uint8_t* __restrict__ ukc, int use_tma,
                                                               int ul_stride /* k-slots per tile in ulist */) {
    constexpr int KP = KROW * 4 + 4;  // padded row stride (doubles), == 4 mod 16
    constexpr int TILE = dense_tile_steps(KROW, TILE_STEPS);  // steps of A staged per buffer
    extern __shared__ __align__(16) double sr_dyn[];  // two buffers of TILE * KP doubles
    __shared__ double sbeta[2][TILE * 2];
    __shared__ double spri[2][TILE];
    __shared__ __align__(8) unsigned long long bar[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;  // group (row of A / column of B), position in the quad
    const int64_t tile = (int64_t)blockIdx.x * 4 + warp;
    const int64_t cbase = tile * (8 * NT);
    // Station compaction (COMPACT, inverse distance): a tile of 8*NT neighbouring cells draws on a dozen of the stations, so the
    // contraction runs over the tile's union list only (ulist: station of k-slot ks*4+q, padded with the index of a zero column;
    // ukc: k-steps in use) -- the zero weights of a 64-station operator are 3 k-steps out of 4.  Otherwise the slots are the stations.
    int ul[KSTEPS];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) ul[ks] = COMPACT ? ulist[tile * ul_stride + ks * 4 + q] : ks * 4 + q;
    const int kc = COMPACT ? min(int(ukc[tile]), KSTEPS) : KSTEPS;
    // B fragments: breg[nt][ks] = omega[cell cbase + nt*8 + g][station of slot ks*4 + q]
    double breg[NT][KSTEPS];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int64_t cell = cbase + nt * 8 + g;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int st = ul[ks];
            breg[nt][ks] = (cell < n_cells && st < n_valid) ? omega[(int64_t)st * n_cells + cell] : 0.0;
        }
    }
    // epilogue constants of the two cells this lane stores per n-tile
    double ez[NT][2], e0[NT][2], e1[NT][2];
    bool eok[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t cell = cbase + nt * 8 + q * 2 + h;
            const bool ok = cell < n_cells && (active == nullptr || active[cell] != 0);
            eok[nt][h] = ok;
            if (MODE == 1) {
                ez[nt][h] = ok ? cz[cell] : 0.0;
                e0[nt][h] = ok ? bm[cell] : 0.0;
                e1[nt][h] = ok ? bm[n_cells + cell] : 0.0;
            } else {
                ez[nt][h] = (ok && bm != nullptr) ? bm[cell] : 0.0;  // addc
                e0[nt][h] = e1[nt][h] = 0.0;
            }
        }
    // both buffers zeroed once: the pad columns (and unused stations) stay zero for the whole launch
    for (int e = threadIdx.x; e < 2 * TILE * KP; e += blockDim.x) sr_dyn[e] = 0.0;
    if (threadIdx.x == 0 && use_tma) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above before the async-proxy copies below
    __syncthreads();
    const int n_tiles = (n_steps + TILE - 1) / TILE;
    const bool vec_ok = (n_cells % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);  // rows of out start 16-byte aligned
    auto stage = [&](int it) {
        const int buf = it & 1, t0 = it * TILE;
        const int nt_steps = min(TILE, n_steps - t0);
        double* dst = sr_dyn + buf * TILE * KP;
        if (use_tma) {
            if (warp == 0) {
                if (lane == 0) mbar_arrive_expect_tx(&bar[buf], (uint32_t)(nt_steps * n_valid * 8));
                __syncwarp();
                for (int r = lane; r < nt_steps; r += 32)
                    bulk_copy_g2s(dst + r * KP, resid + (int64_t)(t0 + r) * row_stride, (uint32_t)(n_valid * 8), &bar[buf]);
            }
        } else {
            for (int e = threadIdx.x; e < TILE * n_valid; e += blockDim.x) {
                const int r = e / n_valid, k = e - r * n_valid;
                dst[r * KP + k] = r < nt_steps ? resid[(int64_t)(t0 + r) * row_stride + k] : 0.0;
            }
        }
        if (MODE == 1)
            for (int e = threadIdx.x; e < TILE; e += blockDim.x) {
                const bool in = e < nt_steps;
                sbeta[buf][2 * e] = in ? beta[2 * (t0 + e)] : 0.0;
                sbeta[buf][2 * e + 1] = in ? beta[2 * (t0 + e) + 1] : 0.0;
                spri[buf][e] = in ? prior_gradient[t0 + e] : 0.0;
            }
    };
    stage(0);
    for (int it = 0; it < n_tiles; ++it) {
        const int buf = it & 1, t0 = it * TILE;
        const int nt_steps = min(TILE, n_steps - t0);
        if (it + 1 < n_tiles) stage(it + 1);  // the other buffer: every warp left it at the barrier that closed iteration it-1
        if (use_tma) mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
        __syncthreads();
        const double* sr = sr_dyn + buf * TILE * KP;
        for (int m0 = 0; m0 < nt_steps; m0 += 8) {
            double acc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                if (!COMPACT || ks < kc) {  // warp-uniform
                    const double a = sr[(m0 + g) * KP + ul[ks]];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(acc[nt][0], acc[nt][1], a, breg[nt][ks]);
                }
            }
            const int tl = m0 + g;  // this lane's row of D
            if (tl < nt_steps) {
                double b0 = 0.0, b1 = 0.0, pri = 0.0;
                if (MODE == 1) { b0 = sbeta[buf][2 * tl]; b1 = sbeta[buf][2 * tl + 1]; pri = spri[buf][tl]; }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double v[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (MODE == 1) {
                            const double t_hat = (1.0 * b0 + ez[nt][h] * b1) + acc[nt][h];
                            v[h] = t_hat - (e0[nt][h] * (b0 - 0.0) + e1[nt][h] * (b1 - pri));
                        } else
                            v[h] = acc[nt][h] + ez[nt][h];
                    }
                    double* dst = out + (int64_t)(t0 + tl) * n_cells + cbase + nt * 8 + q * 2;
                    if (eok[nt][0] && eok[nt][1] && vec_ok) *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);  // the lane's two cells: 16 B
                    else {
                        if (eok[nt][0]) dst[0] = v[0];
                        if (eok[nt][1]) dst[1] = v[1];
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with this buffer before it is refilled
    }
}

// Dense IDW operator of one variable from its neighbour lists, valid while every station value is finite:
//   W[s][c] = w_cs * f_cs / sum_j w_cj  (f = precipitation / radiation factor, 1 otherwise)
//   temperature: + (z_c - zbar_c) / dz * (delta_{s,hi} - delta_{s,lo}) when the neighbours span more than 50 m, else the
//   constant default_gradient * (z_c - zbar_c) goes to addc[c]   (temperature_gradient_scale_computer, inverse_distance.h:304-315)
__global__ void idw_build_dense_kernel(int kind, int64_t n_cells, int n_src, const double* __restrict__ cz, double default_gradient,
                                       const int32_t* __restrict__ nb_idx, const double* __restrict__ nb_w, const double* __restrict__ nb_f,
                                       const int32_t* __restrict__ nb_n, double* __restrict__ W /* [n_src][cells] */, double* __restrict__ addc) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    for (int s = 0; s < n_src; ++s) W[(int64_t)s * n_cells + c] = 0.0;
    const int cnt = nb_n[c];
    addc[c] = 0.0;
    if (cnt == 0) { W[c] = nan_(); return; }  // no station in reach: 0/0 in the reference
    double sum_w = 0.0;
    for (int j = 0; j < cnt; ++j) sum_w += nb_w[(int64_t)j * n_cells + c];
    for (int j = 0; j < cnt; ++j) {
        const int k = nb_idx[(int64_t)j * n_cells + c];
        const double w = nb_w[(int64_t)j * n_cells + c];
        const double f = (kind == IDW_PRECIPITATION || kind == IDW_RADIATION) ? nb_f[(int64_t)j * n_cells + c] : 1.0;
        W[(int64_t)k * n_cells + c] += w * f / sum_w;
    }
    if (kind == IDW_TEMPERATURE) {
        int mn = -1, mx = -1;
        double zmn = 0, zmx = 0, zbar = 0.0;
        for (int j = 0; j < cnt; ++j) {
            const int k = nb_idx[(int64_t)j * n_cells + c];
            const double h = nb_f[(int64_t)j * n_cells + c];  // station height
            zbar += nb_w[(int64_t)j * n_cells + c] * h;
            if (j == 0) { mn = mx = k; zmn = zmx = h; }
            else if (h < zmn) { mn = k; zmn = h; }
            else if (h > zmx) { mx = k; zmx = h; }
        }
        zbar /= sum_w;
        const double dz = zmx - zmn, lever = cz[c] - zbar;
        if (cnt > 1 && dz > 50.0) {
            W[(int64_t)mx * n_cells + c] += lever / dz;
            W[(int64_t)mn * n_cells + c] -= lever / dz;
        } else
            addc[c] = default_gradient * lever;
    }
}

// Union lists for the station compaction of dense_apply_dmma_kernel<KSTEPS, NT, 0>: one warp per tile of 8*NT cells collects, in
// station order, the stations with a non-zero weight for any cell of the tile.
__global__ void idw_union_plan_kernel(int64_t n_cells, int n_src, const double* __restrict__ W /* [n_src][cells] */, int cells_per_tile,
                                      int k_slots, int32_t* __restrict__ ulist /* [tiles][k_slots] */, uint8_t* __restrict__ ukc) {
    const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_cells + cells_per_tile - 1) / cells_per_tile;
    if (tile >= n_tiles) return;  // whole warps leave together
    const int64_t cell = tile * cells_per_tile + lane;
    const bool mine = lane < cells_per_tile && cell < n_cells;
    int count = 0;
    for (int s = 0; s < n_src; ++s) {
        const bool nz = mine && W[(int64_t)s * n_cells + cell] != 0.0;  // NaN (no station in reach) counts as a weight
        if (__any_sync(0xffffffffu, nz)) {
            if (lane == 0) ulist[tile * k_slots + count] = s;
            ++count;
        }
    }
    for (int k = count + lane; k < k_slots; k += 32) ulist[tile * k_slots + k] = k_slots;  // a zero-filled pad column of the staged tile
    if (lane == 0) ukc[tile] = (uint8_t)((count + 3) / 4);
}

__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// identity resampling of average_accessor for a stair-case source on the model axis (time_series.h:202-310): (dt_s*v)/dt_s
__global__ void average_accessor_same_axis_kernel(double* __restrict__ p, int64_t n, double dt_seconds, int* __restrict__ nonfinite) {
    bool bad = false;  // any NaN / inf among the values: the caller keeps a host copy of such series (validity bookkeeping), of no others
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = p[i];
        const bool fin = isfinite(v);
        bad |= !fin;
        p[i] = fin ? (dt_seconds * v) / dt_seconds : nan("");
    }
    if (bad) atomicOr(nonfinite, 1);
}
// average_accessor<point_ts, fixed_dt>::value(i) for sources on their own point axis (core/time_series.h:2033-2072 over
// accumulate_value :202-291): the true average of the source over every model step -- stair-case (POINT_AVERAGE_VALUE) or linear
// between points (POINT_INSTANT_VALUE, strict: nothing after the last point), NaN stretches left out of both area and time,
// NaN at and after the source's total period end (extension policy USE_NAN).  One thread per (model step, source); the left anchor
// is what the reference's hinted search yields under sequential access: the last point at or before the step start, or point 0
// when the source starts later (hint_based_search :165-166 returns 0, not npos, from hint 0).  Same operations in the same order
// as the reference (to_seconds = us / 1e6, a = dv / dt, b = r.v - a * t_r on epoch seconds): bit-identical to the oracle.
__device__ __forceinline__ double us_to_seconds(int64_t us) { return double(us) / 1000000.0; }
__device__ __forceinline__ double average_accessor_value(const int64_t* __restrict__ t /* [n_points] */, const double* __restrict__ values /* point i at values[i * v_stride] */,
                                                         int64_t v_stride, int64_t n_points, int64_t t_end, int linear, int64_t p_start, int64_t p_end) {
    const double nan_v = nan_();
    double result = nan_v;
    if (n_points > 0 && p_start < t_end) {
        int64_t i = 0;
        if (!(p_start < t[0])) {  // index_of(p_start): last point with t <= p_start
            int64_t lo = 0, hi = n_points;  // t[lo] <= p_start < t[hi]
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (t[mid] <= p_start) lo = mid; else hi = mid;
            }
            i = lo;
        }
        const bool extrapolate_flat = !linear;  // strict_linear_between = true
        int64_t l_t = 0, tsum = 0;
        double l_v = 0.0, area = 0.0;
        bool l_finite = false;
        while (true) {
            if (!l_finite) {
                l_t = t[i]; l_v = values[i * v_stride]; ++i;
                l_finite = isfinite(l_v);
                if (i == n_points) {
                    if (l_finite && l_t < p_end && extrapolate_flat) {
                        const int64_t dt = p_end - (p_start > l_t ? p_start : l_t);
                        tsum += dt;
                        area += us_to_seconds(dt) * l_v;
                    }
                    break;
                }
                if (l_t >= p_end) break;
            } else {
                const int64_t r_t = t[i];
                const double r_v = values[i * v_stride];
                ++i;
                const bool r_finite = isfinite(r_v);
                const int64_t px_start = l_t > p_start ? l_t : p_start, px_end = r_t < p_end ? r_t : p_end;
                int64_t dt = px_end - px_start;
                if (linear && r_finite) {
                    const double a = (r_v - l_v) / us_to_seconds(r_t - l_t);
                    const double b = r_v - a * us_to_seconds(r_t);
                    area += us_to_seconds(dt) * (0.5 * a * us_to_seconds(px_start + px_end) + b);
                    tsum += dt;
                } else if (extrapolate_flat) {
                    area += l_v * us_to_seconds(dt);
                    tsum += dt;
                }
                if (i == n_points) {
                    if (r_finite && r_t < p_end && extrapolate_flat) {
                        dt = p_end - r_t;
                        tsum += dt;
                        area += us_to_seconds(dt) * r_v;
                    }
                    break;
                }
                if (r_t >= p_end) break;
                l_finite = r_finite;
                l_t = r_t;
                l_v = r_v;
            }
        }
        result = tsum > 0 ? area / us_to_seconds(tsum) : nan_v;
    }
    return result;
}

__global__ void average_accessor_kernel(const int64_t* __restrict__ t /* [n_points] */, const double* __restrict__ values /* [n_points][n_src] */,
                                        int64_t n_points, int64_t n_src, int64_t t_end, int linear, int64_t ta_t0, int64_t ta_dt, int64_t ta_n,
                                        double* __restrict__ out /* [ta_n][n_src] */) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ta_n * n_src) return;
    const int64_t step = idx / n_src, s = idx - step * n_src;
    const int64_t p_start = ta_t0 + step * ta_dt;
    out[idx] = average_accessor_value(t, values + s, n_src, n_points, t_end, linear, p_start, p_start + ta_dt);
}
// One source projected onto an axis given by its period boundaries (time_axis::point_dt: period i = [pts[i], pts[i+1]))
__global__ void average_accessor_periods_kernel(const int64_t* __restrict__ t, const double* __restrict__ values, int64_t n_points, int64_t t_end,
                                                int linear, const int64_t* __restrict__ pts /* [n + 1] */, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = average_accessor_value(t, values, 1, n_points, t_end, linear, pts[i], pts[i + 1]);
}
// The same for sources that each bring their own point axis (every geo_point_ts of a region_environment is a time-series of its own,
// api/api.h:137-168): source s owns points off[s] .. off[s+1]-1 of the concatenated t / values arrays, its own period end and point
// interpretation.
__global__ void average_accessor_ragged_kernel(const int64_t* __restrict__ t, const double* __restrict__ values, const int64_t* __restrict__ off,
                                               const int64_t* __restrict__ t_end, const int32_t* __restrict__ linear, int64_t n_src, int64_t ta_t0,
                                               int64_t ta_dt, int64_t ta_n, double* __restrict__ out /* [ta_n][n_src] */) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ta_n * n_src) return;
    const int64_t step = idx / n_src, s = idx - step * n_src;
    const int64_t p_start = ta_t0 + step * ta_dt;
    out[idx] = average_accessor_value(t + off[s], values + off[s], 1, off[s + 1] - off[s], t_end[s], linear[s], p_start, p_start + ta_dt);
}

// [rows][cols] -> [cols][rows] tiled transpose (cell-major <-> time-major at the ABI)
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ double tile[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
    }
}
// non-finite values among the calculated cells of a [rows][n_cells] series (region_model.h:954-962)
__global__ void count_nonfinite_kernel(const double* __restrict__ p, int64_t rows, int64_t n_cells, const uint8_t* __restrict__ active,
                                       unsigned long long* __restrict__ count) {
    unsigned long long local = 0;
    const int64_t n = rows * n_cells;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (active != nullptr && active[i % n_cells] == 0) continue;
        local += isfinite(p[i]) ? 0 : 1;
    }
    if (local) atomicAdd(count, local);
}
// single temperature source: its resampled series copied to every calculated cell (region_model.h:470-481)
__global__ void broadcast_source_kernel(int64_t n_cells, const double* __restrict__ src /* [n_steps] */, int n_steps,
                                        const uint8_t* __restrict__ active, double* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || (active != nullptr && active[c] == 0)) return;
    for (int i = 0; i < n_steps; ++i) out[(int64_t)i * n_cells + c] = src[i];
}
// state.adjust_q on the selected cells (region_model.h:831-837)
// cell-identified state io (api/api_state.h:105-140): rows of the [n_state][n_cells] state <-> compact [k][n_state] records
__global__ void state_gather_kernel(const double* __restrict__ state, int64_t n_cells, int n_state, const int64_t* __restrict__ cells, int64_t k,
                                    double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * n_state) return;
    const int64_t r = i / n_state;
    const int s = int(i - r * n_state);
    out[i] = state[(int64_t)s * n_cells + cells[r]];
}
__global__ void state_scatter_kernel(double* __restrict__ state, int64_t n_cells, int n_state, const int64_t* __restrict__ cells, int64_t k,
                                     const double* __restrict__ in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * n_state) return;
    const int64_t r = i / n_state;
    const int s = int(i - r * n_state);
    state[(int64_t)s * n_cells + cells[r]] = in[i];
}
__global__ void scale_selected_kernel(double* __restrict__ v, const uint8_t* __restrict__ sel, int64_t n, double scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && sel[i]) v[i] *= scale;
}

}  // namespace sb2
