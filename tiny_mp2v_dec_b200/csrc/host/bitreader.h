// MSB-first bit reader over a padded byte buffer (64-bit window).
//
// Contract inherited from the reference's decode API: the caller pads the stream with
// >= 64 readable bytes; like the reference's 32-bit word reader (bitstream.h:28-34) this reader may
// touch a few bytes past the last syntax element but never more than 12 past the position it consumed.
//
// Two refill strategies behind one interface (the slice syntax walk in slice_core.h is compiled for
// both):
//   host    one unaligned 8-byte load + bswap per refill, byte granular: refill() leaves >= 56 bits;
//   device  a 64-bit window of two aligned words plus the bit position inside it: skip() is ONE add, peek()
//           a 64-bit shift of the window, refill() a compare (the window slides by a word once the position
//           passes 32, with the load for the word after next already issued -- a one-word look-ahead -- so
//           that no load sits on the dependent chain of the symbol loop; a GPU thread walking a slice is an
//           instruction-count and latency chain); refill() leaves >= 33 bits.
// kBitsAfterRefill is what callers may consume between two refills.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define MP2V_HD __host__ __device__
#else
#define MP2V_HD
#endif

namespace mp2v {

class bitreader_t {
public:
#ifdef __CUDA_ARCH__
    static constexpr int kBitsAfterRefill = 33;
#else
    static constexpr int kBitsAfterRefill = 56;
#endif
    bitreader_t() = default;
    MP2V_HD explicit bitreader_t(const uint8_t* p) { reset(p); }

#ifdef __CUDA_ARCH__
    __device__ void reset(const uint8_t* p) {
        const uintptr_t a = (uintptr_t)p;
        wptr_ = (const uint32_t*)(a & ~(uintptr_t)3);
        hi_ = __byte_perm(__ldg(wptr_), 0u, 0x0123);
        lo_ = __byte_perm(__ldg(wptr_ + 1), 0u, 0x0123);
        wptr_ += 2;
        next_ = __ldg(wptr_);
        pos_ = ((uint32_t)a & 3u) * 8u;                          // drop the bytes before p
    }
    __device__ __forceinline__ void refill() {
        if (pos_ >= 32u) {
            slide();
            if (pos_ >= 32u) slide();                            // only after a gulp of more than 32 bits
        }
    }
#else
    void reset(const uint8_t* p) { ptr_ = p; buf_ = 0; cnt_ = 0; refill(); }
    inline void refill() {
        uint64_t w;
        memcpy(&w, ptr_, 8);
        w = __builtin_bswap64(w);
        buf_ |= w >> cnt_;
        const int adv = (63 - cnt_) >> 3;
        ptr_ += adv;
        cnt_ += adv << 3;
    }
#endif
    // n in 1..32; valid after refill() as long as no more than kBitsAfterRefill bits were consumed since
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(((((uint64_t)hi_ << 32) | lo_) << pos_) >> 32) >> (32 - n); }
    __device__ __forceinline__ void skip(int n) { pos_ += (uint32_t)n; }
    __device__ __forceinline__ uint32_t get(int n) { refill(); const uint32_t v = peek(n); skip(n); return v; }
    __device__ __forceinline__ uint32_t get1() { return get(1); }
#else
    inline uint32_t peek(int n) const { return (uint32_t)(buf_ >> (64 - n)); }
    inline void skip(int n) { buf_ <<= n; cnt_ -= n; }
    inline uint32_t get(int n) { refill(); const uint32_t v = peek(n); skip(n); return v; }
    inline uint32_t get1() { refill(); const uint32_t v = (uint32_t)(buf_ >> 63); skip(1); return v; }
#endif

private:
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ void slide() {
        // the byte swap happens here, at the use: swapping right behind the load would park the
        // (in-order) thread on the load it is supposed to run ahead of
        hi_ = lo_;
        lo_ = __byte_perm(next_, 0u, 0x0123);
        next_ = __ldg(++wptr_);
        pos_ -= 32u;
    }
    const uint32_t* wptr_ = nullptr;   // the word held in next_
    uint32_t next_ = 0;                // look-ahead word as loaded (little-endian view of big-endian data)
    uint32_t hi_ = 0, lo_ = 0;         // the window: 64 bits of the stream, most significant first
    uint32_t pos_ = 0;                 // bits of the window already consumed (< 32 after refill)
#else
    const uint8_t* ptr_ = nullptr;
    uint64_t buf_ = 0;   // unread bits, left aligned
    int cnt_ = 0;        // number of valid bits in buf_
#endif
};

}  // namespace mp2v
