// MSB-first bit reader over a padded byte buffer (64-bit window, branch-light refill).
//
// Contract inherited from the reference's decode API: the caller pads the stream with
// >= 64 readable bytes; like the reference's 32-bit word reader (bitstream.h:28-34) this reader may
// touch a few bytes past the last syntax element but never more than 8 past the position it consumed.
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
    bitreader_t() = default;
    MP2V_HD explicit bitreader_t(const uint8_t* p) { reset(p); }
    MP2V_HD void reset(const uint8_t* p) { ptr_ = p; buf_ = 0; cnt_ = 0; refill(); }

    // make at least 56 bits available
    MP2V_HD inline void refill() {
        uint64_t w;
#ifdef __CUDA_ARCH__
        // device: 8 bytes from an arbitrary address = three aligned words + two byte permutes
        // (reads at most 11 bytes past ptr_; the staged bitstream carries 16 bytes of padding)
        const uintptr_t a = (uintptr_t)ptr_;
        const uint32_t* p32 = (const uint32_t*)(a & ~(uintptr_t)3);
        const uint32_t sel = 0x0123u + 0x1111u * ((uint32_t)a & 3u);
        const uint32_t w0 = __ldg(p32), w1 = __ldg(p32 + 1), w2 = __ldg(p32 + 2);
        w = ((uint64_t)__byte_perm(w0, w1, sel) << 32) | __byte_perm(w1, w2, sel);
#else
        memcpy(&w, ptr_, 8);
        w = __builtin_bswap64(w);
#endif
        buf_ |= w >> cnt_;
        const int adv = (63 - cnt_) >> 3;
        ptr_ += adv;
        cnt_ += adv << 3;
    }
    // n in 1..32; valid after refill() as long as no more than 56 bits were consumed since
    MP2V_HD inline uint32_t peek(int n) const { return (uint32_t)(buf_ >> (64 - n)); }
    MP2V_HD inline uint32_t peek32() const { return (uint32_t)(buf_ >> 32); }
    MP2V_HD inline int bits_left() const { return cnt_; }
    MP2V_HD inline void skip(int n) { buf_ <<= n; cnt_ -= n; }
    MP2V_HD inline uint32_t get(int n) { refill(); const uint32_t v = peek(n); skip(n); return v; }
    MP2V_HD inline uint32_t get1() { refill(); const uint32_t v = (uint32_t)(buf_ >> 63); skip(1); return v; }
    // position of the next unread bit, in bytes from `base` (rounded down)
    MP2V_HD inline const uint8_t* byte_pos() const { return ptr_ - ((cnt_ + 7) >> 3); }

private:
    const uint8_t* ptr_ = nullptr;
    uint64_t buf_ = 0;   // unread bits, left aligned
    int cnt_ = 0;        // number of valid bits in buf_
};

}  // namespace mp2v
