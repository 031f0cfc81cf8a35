// Source-compatibility shim for code that includes the reference's "core/common/cpu.hpp" only for its
// aligned allocator (the sample does: tiny_decoder/tiny_mp2v_dec.cpp:9).  A minimal C++17 allocator
// with the same name and template signature; nothing else of that header is part of the decode API.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <stdlib.h>
#include <new>

template <typename T, std::size_t N = 16>
class AlignmentAllocator {
public:
    using value_type = T;
    template <typename U> struct rebind { using other = AlignmentAllocator<U, N>; };
    AlignmentAllocator() noexcept = default;
    template <typename U> AlignmentAllocator(const AlignmentAllocator<U, N>&) noexcept {}
    T* allocate(std::size_t n) {
        const std::size_t bytes = (n * sizeof(T) + N - 1) / N * N;
        void* p = nullptr;
        if (posix_memalign(&p, N < sizeof(void*) ? sizeof(void*) : N, bytes ? bytes : N) != 0) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, std::size_t) noexcept { std::free(p); }
    template <typename U> bool operator==(const AlignmentAllocator<U, N>&) const noexcept { return true; }
    template <typename U> bool operator!=(const AlignmentAllocator<U, N>&) const noexcept { return false; }
};
