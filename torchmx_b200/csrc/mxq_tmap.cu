// Tensor-map (TMA descriptor) builders of the K3 kernels, with a process-wide cache: see mxq_tc.cuh.
#include <cstring>
#include <mutex>

#include "mxq_tc.cuh"

namespace mxq {
namespace gemm {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(sym);
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

struct Key {
    uint64_t kind;  // 0 operand, 1 scale, 2 output
    uint64_t base;
    int64_t v[5];
    int32_t box_rows, fmt;
};

struct Entry {
    Key key;
    CUtensorMap map;
    bool valid;
};

constexpr int kSlots = 4096;  // direct-mapped; a colliding key simply replaces the slot
Entry g_table[kSlots];
std::mutex g_mu;

uint64_t hash_key(const Key& k) {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < sizeof(Key) / 8; ++i) {
        h ^= w[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xFF51AFD7ED558CCDull;
    }
    return h ^ (h >> 33);
}

template <typename Encode>
bool lookup_or_encode(const Key& key, CUtensorMap* out, Encode&& encode) {
    Entry& e = g_table[hash_key(key) & (kSlots - 1)];
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (e.valid && memcmp(&e.key, &key, sizeof(Key)) == 0) {
            *out = e.map;
            return true;
        }
    }
    if (!encode(out)) return false;
    std::lock_guard<std::mutex> lock(g_mu);
    e.key = key;
    e.map = *out;
    e.valid = true;
    return true;
}

}  // namespace

bool cached_operand_map(CUtensorMap* map, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_rows,
                        int operand_format, int device) {
    (void)device;  // device addresses are unique per process (UVA): the address already identifies the device
    if (batch <= 1) batch_stride = ld * rows;
    Key key;
    memset(&key, 0, sizeof(key));
    key.kind = 0; key.base = (uint64_t)(uintptr_t)base;
    key.v[0] = K; key.v[1] = rows; key.v[2] = batch; key.v[3] = ld; key.v[4] = batch_stride;
    key.box_rows = box_rows; key.fmt = operand_format;
    return lookup_or_encode(key, map, [&](CUtensorMap* m) {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return false;
        const CUtensorMapDataType dt = (operand_format == MXQ_OPERAND_E4M3_BYTES || operand_format == MXQ_OPERAND_E5M2_BYTES) ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                       : (operand_format == MXQ_OPERAND_E2M1_PACKED ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_16U6_ALIGN16B);
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)batch_stride};
        cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return fn(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
}

bool cached_scale_map(CUtensorMap* map, const void* base, int64_t scale_bytes_per_row, int64_t rows, int64_t ld, int device) {
    (void)device;
    Key key;
    memset(&key, 0, sizeof(key));
    key.kind = 1; key.base = (uint64_t)(uintptr_t)base;
    key.v[0] = scale_bytes_per_row; key.v[1] = rows; key.v[2] = ld;
    return lookup_or_encode(key, map, [&](CUtensorMap* m) {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return false;
        cuuint64_t dims[2] = {(cuuint64_t)scale_bytes_per_row, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)ld};
        cuuint32_t box[2] = {16, 128};
        cuuint32_t estr[2] = {1, 1};
        return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
}

bool cached_raw_map(CUtensorMap* map, const void* base, int64_t k_bytes, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_bytes,
                    int box_rows, int device) {
    (void)device;
    if (batch <= 1) batch_stride = ld * rows;
    Key key;
    memset(&key, 0, sizeof(key));
    key.kind = 3; key.base = (uint64_t)(uintptr_t)base;
    key.v[0] = k_bytes; key.v[1] = rows; key.v[2] = batch; key.v[3] = ld; key.v[4] = batch_stride;
    key.box_rows = box_rows; key.fmt = box_bytes;
    return lookup_or_encode(key, map, [&](CUtensorMap* m) {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return false;
        cuuint64_t dims[3] = {(cuuint64_t)k_bytes, (cuuint64_t)rows, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)batch_stride};
        cuuint32_t box[3] = {(cuuint32_t)box_bytes, (cuuint32_t)box_rows, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
}

bool cached_d_map(CUtensorMap* map, void* base, int64_t N, int64_t M, int64_t batch, int64_t ldd, int64_t batch_stride, int device) {
    (void)device;
    if (batch <= 1) batch_stride = ldd * M;
    Key key;
    memset(&key, 0, sizeof(key));
    key.kind = 2; key.base = (uint64_t)(uintptr_t)base;
    key.v[0] = N; key.v[1] = M; key.v[2] = batch; key.v[3] = ldd; key.v[4] = batch_stride;
    return lookup_or_encode(key, map, [&](CUtensorMap* m) {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return false;
        cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)ldd * 2, (cuuint64_t)batch_stride * 2};
        cuuint32_t box[3] = {64, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
}

bool cached_bf16_operand_map(CUtensorMap* map, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_rows,
                             int device) {
    (void)device;
    if (batch <= 1) batch_stride = ld * rows;
    Key key;
    memset(&key, 0, sizeof(key));
    key.kind = 4; key.base = (uint64_t)(uintptr_t)base;
    key.v[0] = K; key.v[1] = rows; key.v[2] = batch; key.v[3] = ld; key.v[4] = batch_stride;
    key.box_rows = box_rows;
    return lookup_or_encode(key, map, [&](CUtensorMap* m) {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return false;
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    });
}

cudaError_t ensure_smem_attr(const void* kernel, int bytes, int device) {
    struct Seen { const void* kernel; uint64_t devices; };
    static Seen seen[128];
    static int n_seen = 0;
    static std::mutex mu;
    const uint64_t bit = (device >= 0 && device < 64) ? (1ull << device) : 0;
    Seen* slot = nullptr;
    if (bit) {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < n_seen; ++i)
            if (seen[i].kernel == kernel) { slot = &seen[i]; break; }
        if (slot && (slot->devices & bit)) return cudaSuccess;
        if (!slot && n_seen < 128) { slot = &seen[n_seen++]; slot->kernel = kernel; slot->devices = 0; }
    }
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && slot) {
        std::lock_guard<std::mutex> lock(mu);
        slot->devices |= bit;
    }
    return e;
}

}  // namespace gemm
}  // namespace mxq
