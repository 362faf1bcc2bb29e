// Index-build helpers next to the path.
//
//   fz_token_starts / fz_hash_tokens   whitespace tokenisation of a UTF-8 byte buffer with Python's str.split() rules
//                                      (src/retrievers/bm25.py:54-60,72,81,101: `doc.split()`), one 128-bit hash per token.
//                                      Equal tokens get equal hashes; the host turns distinct hashes into term ids with
//                                      a sort (torch.unique), so term ids are consistent between corpus and queries.
//   fz_quantiles_f64                   linear-interpolation quantiles of a sorted array: pandas Series.quantile(linspace)
//                                      as used for the percentile distributions (src/retrievers/hybrid.py:391-398).
#include "common.cuh"

namespace fz {

// length in bytes of the whitespace CHARACTER that starts at byte i (0 if none).  Python's str.isspace() set:
// \t \n \v \f \r \x1c-\x1f space | U+0085 U+00A0 | U+1680 U+2000-200A U+2028 U+2029 U+202F U+205F U+3000
__device__ __forceinline__ int ws_len(const unsigned char* __restrict__ b, long long i, long long n) {
    const unsigned c = b[i];
    if (c == 0x20 || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f)) return 1;
    if (c == 0xc2 && i + 1 < n) {
        const unsigned d = b[i + 1];
        return (d == 0x85 || d == 0xa0) ? 2 : 0;
    }
    if (i + 2 < n) {
        const unsigned d = b[i + 1], e = b[i + 2];
        if (c == 0xe1) return (d == 0x9a && e == 0x80) ? 3 : 0;
        if (c == 0xe2) {
            if (d == 0x80 && ((e >= 0x80 && e <= 0x8a) || e == 0xa8 || e == 0xa9 || e == 0xaf)) return 3;
            if (d == 0x81 && e == 0x9f) return 3;
            return 0;
        }
        if (c == 0xe3) return (d == 0x80 && e == 0x80) ? 3 : 0;
    }
    return 0;
}

__device__ __forceinline__ bool is_ws_byte(const unsigned char* __restrict__ b, long long i, long long n) {
    if (ws_len(b, i, n) > 0) return true;
    if (i >= 1 && ws_len(b, i - 1, n) >= 2) return true;
    if (i >= 2 && ws_len(b, i - 2, n) == 3) return true;
    return false;
}

__global__ void token_starts_kernel(const unsigned char* __restrict__ bytes, long long n, unsigned char* __restrict__ flags) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool ws = is_ws_byte(bytes, i, n);
        flags[i] = (!ws && (i == 0 || is_ws_byte(bytes, i - 1, n))) ? 1 : 0;
    }
}

__global__ void hash_tokens_kernel(const unsigned char* __restrict__ bytes, long long n, const int64_t* __restrict__ starts,
                                   long long n_tokens, int64_t* __restrict__ h1, int64_t* __restrict__ h2,
                                   int32_t* __restrict__ lengths) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_tokens; t += (long long)gridDim.x * blockDim.x) {
        long long i = starts[t];
        uint64_t a = 0xcbf29ce484222325ull;                 // FNV-1a 64
        uint64_t b = 0x9e3779b97f4a7c15ull;                 // an independent multiply-xorshift stream
        int len = 0;
        while (i < n && !is_ws_byte(bytes, i, n)) {
            const uint64_t c = bytes[i];
            a = (a ^ c) * 0x100000001b3ull;
            b = (b + c + 1) * 0xff51afd7ed558ccdull;
            b ^= b >> 29;
            ++i;
            ++len;
        }
        h1[t] = (int64_t)a;
        h2[t] = (int64_t)(b ^ ((uint64_t)len << 48));
        if (lengths) lengths[t] = len;
    }
}

// out[i] = np.percentile(sorted, 100 * q_i, method='linear') with q_i = linspace(0, 1, n_q)[i], following numpy step by
// step: q_i = i * (1 / (n_q - 1)) (last = 1), percent = q * 100, quantile = percent / 100, virtual index
// n*q + (1 + q*(-1)) - 1, neighbours floor / floor + 1 (both the last element at or past n - 1), and _lerp:
// a + (b - a) * t, taken from the upper end when t >= 0.5.
__global__ void quantiles_kernel(const double* __restrict__ sorted, long long n, int n_q, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q) return;
    double q = 0.0;
    if (n_q > 1) {
        const double step = __ddiv_rn(1.0, (double)(n_q - 1));
        q = i == n_q - 1 ? 1.0 : __dmul_rn((double)i, step);
    }
    q = __ddiv_rn(__dmul_rn(q, 100.0), 100.0);
    const double nn = (double)n;
    const double virt = __dadd_rn(__dadd_rn(__dmul_rn(nn, q), __dadd_rn(1.0, __dmul_rn(q, -1.0))), -1.0);
    long long lo, hi;
    if (virt >= (double)(n - 1)) { lo = hi = n - 1; }
    else if (virt < 0.0) { lo = hi = 0; }
    else { lo = (long long)floor(virt); hi = lo + 1; }
    const double t = __dadd_rn(virt, -(double)(long long)floor(virt));
    const double a = sorted[lo], b = sorted[hi];
    const double d = __dadd_rn(b, -a);
    out[i] = t >= 0.5 ? __dadd_rn(b, -__dmul_rn(d, __dadd_rn(1.0, -t))) : __dadd_rn(a, __dmul_rn(d, t));
}

}  // namespace fz

using namespace fz;

extern "C" {

int fz_token_starts(const void* utf8, int64_t n_bytes, void* out_flags, fz_stream_t stream) {
    FZ_REQUIRE(n_bytes >= 0 && (n_bytes == 0 || (utf8 && out_flags)), "null pointer");
    if (n_bytes == 0) return FZ_OK;
    const int blocks = (int)(ceil_div<long long>(n_bytes, 256) < 148 * 32 ? ceil_div<long long>(n_bytes, 256) : 148 * 32);
    token_starts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)utf8, n_bytes, (unsigned char*)out_flags);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_hash_tokens(const void* utf8, int64_t n_bytes, const int64_t* starts, int64_t n_tokens, int64_t* out_h1,
                   int64_t* out_h2, int32_t* out_len, fz_stream_t stream) {
    FZ_REQUIRE(n_tokens >= 0 && (n_tokens == 0 || (utf8 && starts && out_h1 && out_h2)), "null pointer");
    if (n_tokens == 0) return FZ_OK;
    const int blocks = (int)(ceil_div<long long>(n_tokens, 256) < 148 * 32 ? ceil_div<long long>(n_tokens, 256) : 148 * 32);
    hash_tokens_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)utf8, n_bytes, starts, n_tokens, out_h1,
                                                                out_h2, out_len);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

int fz_quantiles_f64(const double* sorted, int64_t n, int n_quantiles, double* out, fz_stream_t stream) {
    FZ_REQUIRE(sorted && out && n >= 1 && n_quantiles >= 1, "bad arguments");
    quantiles_kernel<<<ceil_div(n_quantiles, 256), 256, 0, (cudaStream_t)stream>>>(sorted, n, n_quantiles, out);
    FZ_LAUNCH_CHECK();
    return FZ_OK;
}

}  // extern "C"
