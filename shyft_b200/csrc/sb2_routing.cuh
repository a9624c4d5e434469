// sb2_routing.cuh -- cell-to-river unit-hydrograph convolution and river-network accumulation as shared-memory kernels.
//
// Follows core/routing.h:326-383 (routing::model: cell_uhg, cell_output_m3s, local_inflow, upstream_inflow, output_m3s),
// core/routing.h:399-421 (make_uhg_from_gamma, host side in sb2_host.hpp) and the USE_ZERO convolution of
// core/time_series.h:966-975.  The reference recurses upstream per query; here rivers are levelled once on the host
// (leaves first) and every level is one launch, so each river is convolved exactly once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "sb2_host.hpp"

namespace sb2 {

constexpr int ROUTE_CELLS = 128;  // cells per chunk (= threads per block)
constexpr int ROUTE_TT = 64;      // time steps per block

// local_inflow[r][t] = sum over the river's cells (ascending cell order, 128 at a time) of sum_j w_c[j] * q_c[t-j]
// grid (time tiles, rivers); dynamic smem: (ROUTE_TT + max_len - 1) * ROUTE_CELLS doubles
__global__ void __launch_bounds__(ROUTE_CELLS) route_local_inflow_kernel(const double* __restrict__ q /* [T][n_cells] */, int64_t n_cells, int64_t T,
                                                                         const int32_t* __restrict__ riv_ptr, const int32_t* __restrict__ riv_cells,
                                                                         const int32_t* __restrict__ cell_uhg_id /* per gathered cell */,
                                                                         const int32_t* __restrict__ uhg_len, const double* __restrict__ uhg_w,
                                                                         int max_len, double* __restrict__ local /* [n_riv][T] */) {
    extern __shared__ double sq[];  // [(ROUTE_TT + max_len - 1)][ROUTE_CELLS]
    __shared__ double acc[ROUTE_TT];
    const int r = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * ROUTE_TT;
    const int nt = int(min((int64_t)ROUTE_TT, T - t0));
    const int rows = ROUTE_TT + max_len - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < ROUTE_TT) acc[threadIdx.x] = 0.0;
    for (int k0 = riv_ptr[r]; k0 < riv_ptr[r + 1]; k0 += ROUTE_CELLS) {
        const int nc = min(ROUTE_CELLS, riv_ptr[r + 1] - k0);
        __syncthreads();
        // stage q rows [t0 - (max_len-1), t0 + ROUTE_TT) of this chunk's cells; rows before the axis start are zero (USE_ZERO)
        {
            const int c = threadIdx.x;
            const int64_t cell = c < nc ? riv_cells[k0 + c] : -1;
            for (int row = 0; row < rows; ++row) {
                const int64_t t = t0 - (max_len - 1) + row;
                sq[row * ROUTE_CELLS + c] = (cell >= 0 && t >= 0 && t < T) ? q[t * n_cells + cell] : 0.0;
            }
        }
        __syncthreads();
        // warp w reduces rows t = w, w+4, ...; lane l covers cells l, l+32, l+64, l+96
        for (int ti = warp; ti < nt; ti += ROUTE_CELLS / 32) {
            double s = 0.0;
            for (int c = lane; c < nc; c += 32) {
                const int id = cell_uhg_id[k0 + c];
                const int len = uhg_len[id];
                const double* w = uhg_w + (int64_t)id * max_len;
                double v = 0.0;
                for (int j = 0; j < len; ++j) v += w[j] * sq[(ti + max_len - 1 - j) * ROUTE_CELLS + c];
                s += v;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) acc[ti] += s;
        }
    }
    __syncthreads();
    if (threadIdx.x < nt) local[(int64_t)r * T + t0 + threadIdx.x] = acc[threadIdx.x];
}

// one network level: upstream[r][t] = sum of output[u][t] over upstream rivers u; output[r] = (local[r] + upstream[r]) (*) uhg_r
// grid (time tiles, rivers of this level); dynamic smem: (ROUTE_TT + max_len - 1) doubles
__global__ void __launch_bounds__(ROUTE_TT) route_river_level_kernel(const int32_t* __restrict__ level_rivers, const int32_t* __restrict__ up_ptr,
                                                                     const int32_t* __restrict__ up_idx, const int32_t* __restrict__ riv_uhg_id,
                                                                     const int32_t* __restrict__ uhg_len, const double* __restrict__ uhg_w, int max_len,
                                                                     int64_t T, const double* __restrict__ local, double* __restrict__ upstream,
                                                                     double* __restrict__ output) {
    extern __shared__ double sin_[];  // input sum rows [t0 - (max_len-1), t0 + ROUTE_TT)
    const int r = level_rivers[blockIdx.y];
    const int64_t t0 = (int64_t)blockIdx.x * ROUTE_TT;
    const int rows = ROUTE_TT + max_len - 1;
    for (int row = threadIdx.x; row < rows; row += blockDim.x) {
        const int64_t t = t0 - (max_len - 1) + row;
        double v = 0.0;
        if (t >= 0 && t < T) {
            double up = 0.0;
            for (int k = up_ptr[r]; k < up_ptr[r + 1]; ++k) up += output[(int64_t)up_idx[k] * T + t];
            if (row >= max_len - 1) upstream[(int64_t)r * T + t] = up;
            v = local[(int64_t)r * T + t] + up;
        }
        sin_[row] = v;
    }
    __syncthreads();
    const int64_t t = t0 + threadIdx.x;
    if (t < T) {
        const int id = riv_uhg_id[r];
        const int len = uhg_len[id];
        const double* w = uhg_w + (int64_t)id * max_len;
        double v = 0.0;
        for (int j = 0; j < len; ++j) v += w[j] * sin_[threadIdx.x + max_len - 1 - j];
        output[(int64_t)r * T + t] = v;
    }
}

namespace routing_host {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T>
    T* upload(const std::vector<T>& h, cudaStream_t s) {
        if (h.empty()) return nullptr;
        if (cudaMalloc(&p, h.size() * sizeof(T)) != cudaSuccess) throw std::runtime_error("routing: device allocation failed");
        if (cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess)
            throw std::runtime_error("routing: upload failed");
        return (T*)p;
    }
    template <class T>
    T* alloc(size_t n) {
        if (cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)) != cudaSuccess) throw std::runtime_error("routing: device allocation failed");
        return (T*)p;
    }
};

struct UhgTable {
    std::map<std::tuple<int, double, double>, int> ids;
    std::vector<std::vector<double>> w;
    int id_of(int n_steps, double alpha, double beta) {
        auto key = std::make_tuple(n_steps, alpha, beta);
        auto f = ids.find(key);
        if (f != ids.end()) return f->second;
        w.push_back(host::make_uhg_from_gamma(n_steps, alpha, beta));
        return ids[key] = int(w.size()) - 1;
    }
};

}  // namespace routing_host

// cell_routing [n_cells][5] = routing id, distance, velocity, alpha, beta (velocity/alpha/beta from the cell's parameter set);
// rivers [n][6] = id, downstream id, distance, velocity, alpha, beta.  Writes the three series of river `rid` for
// [start_step, start_step + n_steps) to host memory (any of them may be null).
inline void route_rivers(const std::vector<double>& rivers, int64_t rid, const std::vector<double>& cell_routing, int64_t n_cells, int64_t T,
                         int64_t dt_us, const double* d_q, cudaStream_t stream, int64_t* launches, int64_t start_step, int64_t n_steps,
                         double* local_inflow, double* upstream_inflow, double* output) {
    using namespace routing_host;
    const int n_riv = int(rivers.size() / 6);
    if (n_riv == 0) throw std::runtime_error("routing: the river network is empty");
    std::map<int64_t, int> ix_of_rid;
    for (int i = 0; i < n_riv; ++i) ix_of_rid[int64_t(rivers[6 * i])] = i;
    if (!ix_of_rid.count(rid)) throw std::runtime_error("river network: river id " + std::to_string(rid) + " not found");
    UhgTable tab;
    // cells gathered per river, ascending cell order
    std::vector<std::vector<int32_t>> cells_of(n_riv);
    std::vector<int32_t> cell_uhg(n_cells, 0);
    for (int64_t c = 0; c < n_cells; ++c) {
        const int64_t id = int64_t(cell_routing[5 * c]);
        if (id <= 0) continue;
        auto f = ix_of_rid.find(id);
        if (f == ix_of_rid.end()) throw std::runtime_error("river network: river id " + std::to_string(id) + " not found");
        cells_of[f->second].push_back(int32_t(c));
        cell_uhg[c] = tab.id_of(host::uhg_steps(cell_routing[5 * c + 1], cell_routing[5 * c + 2], dt_us), cell_routing[5 * c + 3],
                                cell_routing[5 * c + 4]);
    }
    std::vector<int32_t> riv_ptr(n_riv + 1, 0), riv_cells, gathered_uhg, riv_uhg(n_riv);
    for (int r = 0; r < n_riv; ++r) {
        for (auto c : cells_of[r]) { riv_cells.push_back(c); gathered_uhg.push_back(cell_uhg[c]); }
        riv_ptr[r + 1] = int32_t(riv_cells.size());
        riv_uhg[r] = tab.id_of(host::uhg_steps(rivers[6 * r + 2], rivers[6 * r + 3], dt_us), rivers[6 * r + 4], rivers[6 * r + 5]);
    }
    int max_len = 1;
    for (auto& w : tab.w) max_len = std::max(max_len, int(w.size()));
    const size_t smem_cells = size_t(ROUTE_TT + max_len - 1) * ROUTE_CELLS * sizeof(double);
    if (smem_cells > 200 * 1024) throw std::runtime_error("routing: unit hydrograph too long for the shared-memory tile");
    std::vector<int32_t> uhg_len(tab.w.size());
    std::vector<double> uhg_w(tab.w.size() * max_len, 0.0);
    for (size_t i = 0; i < tab.w.size(); ++i) {
        uhg_len[i] = int32_t(tab.w[i].size());
        std::copy(tab.w[i].begin(), tab.w[i].end(), uhg_w.begin() + i * max_len);
    }
    // upstream lists (ascending river id, routing.h:205-213) and levels (a river's level = 1 + max level of its upstreams)
    std::vector<std::vector<int32_t>> ups(n_riv);
    for (auto& kv : ix_of_rid) {  // map order = ascending id
        const int64_t down = int64_t(rivers[6 * kv.second + 1]);
        if (down > 0) {
            auto f = ix_of_rid.find(down);
            if (f == ix_of_rid.end()) throw std::runtime_error("river network: downstream river id not found");
            ups[f->second].push_back(kv.second);
        }
    }
    std::vector<int> level(n_riv, -1);
    int n_levels = 0;
    for (int pass = 0; pass <= n_riv; ++pass) {
        bool changed = false;
        for (int r = 0; r < n_riv; ++r) {
            if (level[r] >= 0) continue;
            int lv = 0;
            bool ready = true;
            for (auto u : ups[r]) { if (level[u] < 0) { ready = false; break; } lv = std::max(lv, level[u] + 1); }
            if (ready) { level[r] = lv; n_levels = std::max(n_levels, lv + 1); changed = true; }
        }
        if (!changed) break;
    }
    for (int r = 0; r < n_riv; ++r)
        if (level[r] < 0) throw std::runtime_error("river network: cycle detected");
    std::vector<int32_t> up_ptr(n_riv + 1, 0), up_idx;
    for (int r = 0; r < n_riv; ++r) { for (auto u : ups[r]) up_idx.push_back(u); up_ptr[r + 1] = int32_t(up_idx.size()); }

    DevBuf b_ptr, b_cells, b_cuhg, b_len, b_w, b_ruhg, b_upp, b_upi, b_local, b_up, b_out, b_lvl;
    const int32_t* d_ptr = b_ptr.upload(riv_ptr, stream);
    const int32_t* d_cells = b_cells.upload(riv_cells, stream);
    const int32_t* d_cuhg = b_cuhg.upload(gathered_uhg, stream);
    const int32_t* d_len = b_len.upload(uhg_len, stream);
    const double* d_w = b_w.upload(uhg_w, stream);
    const int32_t* d_ruhg = b_ruhg.upload(riv_uhg, stream);
    const int32_t* d_upp = b_upp.upload(up_ptr, stream);
    const int32_t* d_upi = b_upi.upload(up_idx, stream);
    double* d_local = b_local.alloc<double>(size_t(n_riv) * T);
    double* d_up = b_up.alloc<double>(size_t(n_riv) * T);
    double* d_out = b_out.alloc<double>(size_t(n_riv) * T);
    cudaMemsetAsync(d_up, 0, size_t(n_riv) * T * sizeof(double), stream);

    const unsigned tiles = unsigned((T + ROUTE_TT - 1) / ROUTE_TT);
    if (smem_cells > 48 * 1024) cudaFuncSetAttribute(route_local_inflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cells));
    route_local_inflow_kernel<<<dim3(tiles, n_riv), ROUTE_CELLS, smem_cells, stream>>>(d_q, n_cells, T, d_ptr, d_cells, d_cuhg, d_len, d_w, max_len,
                                                                                     d_local);
    ++*launches;
    std::vector<int32_t> order;
    std::vector<int> level_begin(n_levels + 1, 0);
    for (int lv = 0; lv < n_levels; ++lv) {
        for (int r = 0; r < n_riv; ++r) if (level[r] == lv) order.push_back(r);
        level_begin[lv + 1] = int(order.size());
    }
    const int32_t* d_order = b_lvl.upload(order, stream);
    const size_t smem_riv = size_t(ROUTE_TT + max_len - 1) * sizeof(double);
    for (int lv = 0; lv < n_levels; ++lv) {
        const int cnt = level_begin[lv + 1] - level_begin[lv];
        route_river_level_kernel<<<dim3(tiles, cnt), ROUTE_TT, smem_riv, stream>>>(d_order + level_begin[lv], d_upp, d_upi, d_ruhg, d_len, d_w, max_len,
                                                                                   T, d_local, d_up, d_out);
        ++*launches;
    }
    if (cudaGetLastError() != cudaSuccess) throw std::runtime_error("routing: kernel launch failed");
    const int r = ix_of_rid[rid];
    const size_t bytes = size_t(n_steps) * sizeof(double);
    if (local_inflow) cudaMemcpyAsync(local_inflow, d_local + size_t(r) * T + start_step, bytes, cudaMemcpyDeviceToHost, stream);
    if (upstream_inflow) cudaMemcpyAsync(upstream_inflow, d_up + size_t(r) * T + start_step, bytes, cudaMemcpyDeviceToHost, stream);
    if (output) cudaMemcpyAsync(output, d_out + size_t(r) * T + start_step, bytes, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) throw std::runtime_error(std::string("routing: ") + cudaGetErrorString(e));
}

}  // namespace sb2
