// CTA-wide radix select: the k-th largest of n composite keys (hi : lo) held in shared memory, most significant
// byte first.  Keys are unique when `lo` carries a unique tie-breaker (~doc id), so "key >= k-th key" selects
// exactly k records.  Every pass histograms one byte of the keys that still match the decided prefix (8-12 cheap
// passes over <= 8192 records) instead of sorting all records.
#pragma once

#include <cstdint>

namespace fz {

template <typename HiT>
struct KeyBits;
template <>
struct KeyBits<uint32_t> { static constexpr int hi_bits = 32; };
template <>
struct KeyBits<uint64_t> { static constexpr int hi_bits = 64; };

// hist: 256 counters in shared memory; bcast: 4 ints in shared memory.  All threads of the CTA must call.
// Requires 1 <= k <= n.  On return (kth_hi, kth_lo) is the k-th largest key.
// Scores of one query share their sign and exponent bits, so whole passes fall into ONE bin: a warp whose keys share
// the digit adds its count with one atomic, and the passes stop as soon as the k-th key's bin holds a single key
// (normally once the score bits are consumed - the doc-id bits only break exact ties).
template <typename HiT, typename GetHi, typename GetLo>
__device__ void cta_radix_select_kth_by(GetHi get_hi, GetLo get_lo, int n, int k, int* hist, int* bcast, HiT& kth_hi,
                                        uint32_t& kth_lo) {
    constexpr int HB = KeyBits<HiT>::hi_bits;
    HiT p_hi = 0, m_hi = 0;          // decided prefix bits / their mask, hi word
    uint32_t p_lo = 0, m_lo = 0;     // same, lo word
    int remaining = k;
    const int lane = threadIdx.x & 31;
    for (int shift = HB + 32 - 8; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const bool in_hi = shift >= 32;
        for (int base = 0; base < n; base += blockDim.x) {
            const int i = base + threadIdx.x;
            bool match = false;
            uint32_t d = 0;
            if (i < n) {
                const HiT h = get_hi(i);
                const uint32_t l = get_lo(i);
                match = (h & m_hi) == p_hi && (l & m_lo) == p_lo;
                d = in_hi ? (uint32_t)(h >> (shift - 32)) & 255u : (l >> shift) & 255u;
            }
            // One shared-memory atomic per warp when all matching lanes share the digit (whole passes do: the scores of
            // one query share sign and exponent), plain atomics on the spread-out digits otherwise.  (__match_any_sync
            // would aggregate every case but costs far more than the conflicts it saves.)
            const unsigned act = __ballot_sync(0xffffffffu, match);
            if (act) {
                const int leader = __ffs(act) - 1;
                const uint32_t d0 = __shfl_sync(0xffffffffu, d, leader);
                const bool uniform = __all_sync(0xffffffffu, !match || d == d0);
                if (uniform) {
                    if (lane == leader) atomicAdd(&hist[d0], __popc(act));
                } else if (match) {
                    atomicAdd(&hist[d], 1);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // lane l owns bins 255-8l .. 248-8l (descending); find the bin where the running count reaches `remaining`
            int c[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; s += c[j]; }
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - s;
            if (excl < remaining && remaining <= incl) {
                int run = excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (run < remaining && remaining <= run + c[j]) {
                        bcast[0] = 255 - 8 * lane - j;      // the digit of the k-th key
                        bcast[1] = remaining - run;         // rank inside that bin
                        bcast[2] = c[j];                    // keys in that bin
                    }
                    run += c[j];
                }
            }
        }
        __syncthreads();
        const uint32_t d = (uint32_t)bcast[0];
        remaining = bcast[1];
        const int in_bin = bcast[2];
        if (in_hi) {
            p_hi |= (HiT)d << (shift - 32);
            m_hi |= (HiT)255 << (shift - 32);
        } else {
            p_lo |= d << shift;
            m_lo |= 255u << shift;
        }
        __syncthreads();
        if (in_bin == 1 && shift > 0) {
            // the k-th key is the only key with this prefix: fetch it instead of histogramming its remaining bytes
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const HiT h = get_hi(i);
                const uint32_t l = get_lo(i);
                if ((h & m_hi) == p_hi && (l & m_lo) == p_lo) {
                    bcast[3] = i;
                }
            }
            __syncthreads();
            const int idx = bcast[3];
            kth_hi = get_hi(idx);
            kth_lo = get_lo(idx);
            __syncthreads();
            return;
        }
    }
    kth_hi = p_hi;
    kth_lo = p_lo;
}

template <typename HiT>
__device__ void cta_radix_select_kth(const HiT* __restrict__ hi, const uint32_t* __restrict__ lo, int n, int k,
                                     int* hist, int* bcast, HiT& kth_hi, uint32_t& kth_lo) {
    cta_radix_select_kth_by<HiT>([hi](int i) { return hi[i]; }, [lo](int i) { return lo[i]; }, n, k, hist, bcast, kth_hi,
                                 kth_lo);
}

template <typename HiT>
__device__ __forceinline__ bool key_ge(HiT a_hi, uint32_t a_lo, HiT b_hi, uint32_t b_lo) {
    return a_hi > b_hi || (a_hi == b_hi && a_lo >= b_lo);
}

}  // namespace fz
