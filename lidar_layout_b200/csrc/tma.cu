// TMA descriptor (CUtensorMap) construction.  cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda.
#include <atomic>

#include "common.h"

namespace lidm {

// Programmatic dependent launch: on for small batches (the per-kernel launch latency is what bounds them; B <= 16 measured
// neutral to +1 %), off for large ones - at B = 64 the dependents' early CTAs compete with the last wave of the multi-wave
// kernels: the power-capped bench loop runs 98.7-99.1 samples/s with it and 100.0-100.1 without, same box
// (profiles/r02j_sustained_ab_pdl.txt).  LIDM_NO_PDL: never; LIDM_PDL_ALWAYS: at every batch size.
static std::atomic<int> g_pdl_large_batch{0};
void pdl_set_batch(int batch) { g_pdl_large_batch.store(batch > 16 ? 1 : 0, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = getenv("LIDM_NO_PDL") == nullptr;
  static const bool always = getenv("LIDM_PDL_ALWAYS") != nullptr;
  return on && (always || g_pdl_large_batch.load(std::memory_order_relaxed) == 0);
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LIDM_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr)
      throw Error(-2, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

CUtensorMapSwizzle swz(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

CUtensorMap encode(const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, int swizzle_bytes, const cuuint32_t* elem_strides = nullptr) {
  CUtensorMap m;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank && elem_strides != nullptr; ++i) estr[i] = elem_strides[i];
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                           strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes),
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string msg = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " rank " +
                      std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string((unsigned long long)dims[i]);
    msg += " strides";
    for (int i = 0; i + 1 < rank; ++i) msg += " " + std::to_string((unsigned long long)strides_bytes[i]);
    msg += " box";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(box[i]);
    throw Error(-2, msg);
  }
  return m;
}
}  // namespace

// 4-D map over a channels-last activation view: dims (C, Wp, H, B), box (box_c, box_w, box_h, 1).
// sx, sy > 1: traversal strides (a strided convolution's operand): the box still delivers box_w x box_h pixels, taken
// every sx-th column / sy-th row from the start coordinate (the driver wants boxDim = pixels * stride there).
CUtensorMap make_tma_act(const View& v, int box_c, int box_w, int box_h, int swizzle_bytes, int box_b, int sx, int sy) {
  cuuint64_t dims[4] = {(cuuint64_t)v.Cphys(), (cuuint64_t)v.Wp(), (cuuint64_t)v.H, (cuuint64_t)v.B};
  cuuint64_t str[3] = {(cuuint64_t)v.ld * 2, (cuuint64_t)v.ld * 2 * v.pitch(), (cuuint64_t)v.ld * 2 * v.pitch() * v.H};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * sx), (cuuint32_t)(box_h * sy), (cuuint32_t)box_b};
  cuuint32_t est[4] = {1, (cuuint32_t)sx, (cuuint32_t)sy, 1};
  LIDM_REQUIRE(sx >= 1 && sy >= 1 && sx <= 8 && sy <= 8 && box_w * sx <= 256 && box_h * sy <= 256, "TMA traversal stride / box");
  return encode(v.p, 4, dims, str, box, swizzle_bytes, (sx > 1 || sy > 1) ? est : nullptr);
}

CUtensorMap make_tma_3d(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                        uint64_t stride2_bytes, uint32_t b0, uint32_t b1, int swizzle_bytes) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t str[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  return encode(base, 3, dims, str, box, swizzle_bytes);
}

}  // namespace lidm
