// Small device-side helpers shared by the kernels of this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>

namespace orbx {

void set_error(const std::string& s);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the (function, device) pair, not of a handle: several handles of
// different sizes share it.  Keep a process-wide monotonic maximum per kernel and device and only ever raise it, so that a
// smaller handle created later can never lower the limit under a larger one (the launch would fail with invalid-value).
template <class Kernel>
inline int ensure_dynamic_smem(Kernel* kernel, size_t bytes, int device) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> current;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = current[std::make_pair((const void*)kernel, device)];
    if (bytes <= cur || bytes <= 48 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e)); return -4; }
    cur = bytes;
    return 0;
}

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan of an int array living in shared memory (in place).
// Returns the total.  `warpTmp` needs 33 ints.  All threads of the block must call.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_excl_scan(int* data, int m, int* warpTmp) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int carry = 0;
    for (int base = 0; base < m; base += nt) {
        const int i = base + tid;
        const int v = i < m ? data[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warpTmp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = lane < nw ? warpTmp[lane] : 0;
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warpTmp[lane] = winc - w;       // exclusive warp offsets
            if (lane == 31) warpTmp[32] = winc;  // tile total
        }
        __syncthreads();
        if (i < m) data[i] = carry + warpTmp[wid] + inc - v;
        carry += warpTmp[32];
        __syncthreads();
    }
    return carry;
}


// 256-bit Hamming distance of two 32-byte descriptors held as 8 words (ORBmatcher::DescriptorDistance,
// reference src/ORBmatcher.cc:2058-2074; the SWAR popcount there == __popc).
__device__ __forceinline__ int hamming256(const uint32_t* a, const uint32_t* b) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d += __popc(a[i] ^ b[i]);
    return d;
}

// Wait for a stream from a host thread that has nothing else to do: through an event created with cudaEventBlockingSync the thread sleeps
// instead of spinning in cudaStreamSynchronize (a server runs several such threads per GPU and several GPUs per host; spinning threads
// take the cores the other ranks' threads need).  Used by the BATCH host entry points; single-frame calls keep the low-latency spin.
inline cudaError_t wait_stream_blocking(cudaStream_t st, cudaEvent_t ev) {
    cudaError_t e = cudaEventRecord(ev, st);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}
inline cudaError_t make_blocking_event(cudaEvent_t* ev) { return cudaEventCreateWithFlags(ev, cudaEventBlockingSync | cudaEventDisableTiming); }
}  // namespace orbx
