// Error reporting, launch accounting and the top-K-table -> scipy-canonical CSR assembly of libb200rec.so.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>

#include <string.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

// entry e of column c (e < cnt[c]) -> compacted position off[c] + e
__global__ void compact_table_kernel(int n_cols, int K, const int* __restrict__ idx, const float* __restrict__ val,
                                     const int* __restrict__ cnt, const int* __restrict__ off, int* rows, int* pos,
                                     int* cols, float* vals, int* row_cnt) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= (long long)n_cols * K) return;
  const int c = (int)(g / K), e = (int)(g % K);
  if (e >= cnt[c]) return;
  const int o = off[c] + e;
  const int r = idx[g];
  rows[o] = r;
  pos[o] = o;
  cols[o] = c;
  vals[o] = val[g];
  atomicAdd(row_cnt + r, 1);
}

__global__ void gather_sorted_kernel(long long nnz, const int* __restrict__ perm, const int* __restrict__ cols,
                                     const float* __restrict__ vals, int* out_cols, float* out_vals) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= nnz) return;
  const int p = perm[g];
  out_cols[g] = cols[p];
  out_vals[g] = vals[p];
}

// Within one column the slots are in selection order; the stable sort by row keeps the (column-major) input
// order inside a row, so compacted entries must be ordered by column: they are, because off[] is the
// exclusive scan over columns.

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return g_err; }
int b200_version(void) { return 100; }
int64_t b200_launch_count(void) { return g_launches.load(); }

int b200_device_info(char* name, int name_len, int* sm_count_out, int64_t* total_mem) {
  return guarded([&] {
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (name && name_len > 0) {
      strncpy(name, prop.name, (size_t)name_len - 1);
      name[name_len - 1] = 0;
    }
    if (sm_count_out) *sm_count_out = prop.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)prop.totalGlobalMem;
  });
}

int b200_topk_table_to_csr_count(int n_cols, int K, const int32_t* d_cnt, int64_t* nnz_out, void* stream) {
  return guarded([&] {
    B200_REQUIRE(n_cols > 0 && K > 0 && d_cnt && nnz_out, "b200_topk_table_to_csr_count: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf<long long> total(1);
    size_t tmp_bytes = 0;
    B200_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, d_cnt, total.get(), n_cols, st));
    DevBuf<unsigned char> tmp(tmp_bytes + 16);
    B200_CUDA(cub::DeviceReduce::Sum(tmp.get(), tmp_bytes, d_cnt, total.get(), n_cols, st));
    count_launch(2);
    long long h = 0;
    B200_CUDA(cudaMemcpyAsync(&h, total.get(), sizeof(long long), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *nnz_out = h;
  });
}

int b200_topk_table_to_csr_fill(int n_cols, int K, const int32_t* d_idx, const float* d_val, const int32_t* d_cnt,
                                int64_t nnz, int32_t* h_indptr, int32_t* h_indices, float* h_data, void* stream) {
  return guarded([&] {
    B200_REQUIRE(n_cols > 0 && K > 0 && d_idx && d_val && d_cnt && h_indptr, "b200_topk_table_to_csr_fill: bad argument");
    B200_REQUIRE(nnz >= 0 && nnz < (1ll << 31) - 1, "b200_topk_table_to_csr_fill: nnz out of int32 range");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n1 = (size_t)std::max<long long>(nnz, 1);
    DevBuf<int> off((size_t)n_cols + 1), row_cnt((size_t)n_cols + 1), indptr((size_t)n_cols + 1);
    DevBuf<int> rows(n1), pos(n1), cols(n1), rows_sorted(n1), perm(n1), out_cols(n1);
    DevBuf<float> vals(n1), out_vals(n1);
    B200_CUDA(cudaMemsetAsync(off.get(), 0, sizeof(int) * ((size_t)n_cols + 1), st));
    B200_CUDA(cudaMemsetAsync(row_cnt.get(), 0, sizeof(int) * ((size_t)n_cols + 1), st));
    B200_CUDA(cudaMemsetAsync(indptr.get(), 0, sizeof(int) * ((size_t)n_cols + 1), st));
    size_t tmp_bytes = 0, tb2 = 0;
    B200_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_cnt, off.get() + 1, n_cols, st));
    int end_bit = 1;
    while ((1ll << end_bit) < (long long)n_cols) ++end_bit;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, rows.get(), rows_sorted.get(), pos.get(), perm.get(), (int)nnz, 0,
                                              end_bit, st));
    DevBuf<unsigned char> tmp(std::max(tmp_bytes, tb2) + 16);
    B200_CUDA(cub::DeviceScan::InclusiveSum(tmp.get(), tmp_bytes, d_cnt, off.get() + 1, n_cols, st));
    count_launch(2);
    if (nnz > 0) {
      compact_table_kernel<<<div_up((long long)n_cols * K, 256), 256, 0, st>>>(n_cols, K, d_idx, d_val, d_cnt, off.get(), rows.get(),
                                                                              pos.get(), cols.get(), vals.get(), row_cnt.get());
      count_launch();
      B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tb2, rows.get(), rows_sorted.get(), pos.get(), perm.get(), (int)nnz, 0,
                                                end_bit, st));
      count_launch(4);
      gather_sorted_kernel<<<div_up(nnz, 256), 256, 0, st>>>(nnz, perm.get(), cols.get(), vals.get(), out_cols.get(), out_vals.get());
      count_launch();
    }
    size_t tb3 = tmp_bytes;
    B200_CUDA(cub::DeviceScan::InclusiveSum(tmp.get(), tb3, row_cnt.get(), indptr.get() + 1, n_cols, st));
    count_launch(2);
    B200_CUDA(cudaMemcpyAsync(h_indptr, indptr.get(), sizeof(int) * ((size_t)n_cols + 1), cudaMemcpyDeviceToHost, st));
    if (nnz > 0) {
      B200_REQUIRE(h_indices && h_data, "b200_topk_table_to_csr_fill: NULL output");
      B200_CUDA(cudaMemcpyAsync(h_indices, out_cols.get(), sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaMemcpyAsync(h_data, out_vals.get(), sizeof(float) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
