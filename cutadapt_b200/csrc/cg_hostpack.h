// cg_hostpack.h -- host side of the compressed host-to-device transfer of cg_process_batch.
//
// The reads of a chunk travel over PCIe as a base-6 stream, three characters per byte
// (A C G T N + "escape"); every other byte value is sent verbatim in an exception list
// (position, byte).  The device expands the stream back into the caller's bytes
// (cg_unpack3_kernel + cg_unpack_fix_kernel), so everything downstream sees exactly the
// caller's input: the transfer is lossless and changes no result.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// A small pool of worker threads for the per-chunk host work of cg_process_batch
// (packing, scanning the offsets).  run() is a blocking parallel-for over job numbers;
// the calling thread takes part.
class CgHostPool {
public:
    explicit CgHostPool(int n_threads);
    ~CgHostPool();
    int size() const { return (int)workers_.size() + 1; }
    void run(int64_t n_jobs, const std::function<void(int64_t job, int worker)> &fn);
    // Run the workers on the CPUs of the NUMA node that holds `addr` (the caller's read buffer): the packer is
    // bound by memory latency, and remote reads cost about half of a thread's bandwidth.  Best effort (Linux,
    // more than one node, get_mempolicy permitted); done once per pool.  Returns the node or -1.
    int follow_memory(const void *addr);

private:
    int followed_node_ = -2;   // -2: not tried yet
    // Workers spin on `generation_` for a while after a job set (the calls of one cg_process_batch come
    // back to back) and only then go to sleep on the condition variable: waking 60 sleeping threads
    // through a mutex costs more than packing a chunk.
    void worker_main(int id);
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_start_;
    const std::function<void(int64_t, int)> *fn_ = nullptr;
    std::atomic<int64_t> next_{0};
    std::atomic<int64_t> n_jobs_{0};
    std::atomic<uint64_t> generation_{0};
    std::atomic<int> pending_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<bool> stop_{false};
};

int cg_host_cpus();              // usable CPUs (affinity mask, cgroup quota)
int cg_host_threads_default();

// Packed stream geometry: stream byte i holds the characters at absolute positions
// a0 + 3i .. a0 + 3i + 2 of the caller's array.
#define CG_PACK_ESCAPE 5

// Pack stream bytes [i0, i1) of the stream that starts at absolute position a0.
// Characters at absolute positions outside [lo, hi) are not read (filler 'A').
// Exceptions are appended to exc as (position relative to a0) << 8 | byte.
void cg_pack3_range(const uint8_t *seq, int64_t a0, int64_t lo, int64_t hi, int64_t i0, int64_t i1,
                    uint8_t *packed, std::vector<uint64_t> &exc);
