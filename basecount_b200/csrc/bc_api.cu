// bc_api.cu -- the C ABI of include/basecount_b200.h over the sm_100a kernels.
//
// Host side of the hot path: owns the device accumulators (the reference's
// np.zeros((L, 6)) at basecount/main.py:132), double-buffered staging for batches that
// arrive in pinned host memory (H2D on a copy stream, kernels on a compute stream), and
// the result read-back.  No CPU compute path exists here: every entry point that
// produces numbers launches CUDA kernels or fails.
#include "../../include/basecount_b200.h"
#include "bc_common.cuh"
#include "k1_count.cuh"
#include "k1_fast.cuh"
#include "k2_stats.cuh"
#include "k3_reduce.cuh"
#include "bam_decode.h"
#include "bam_index.h"
#include "tsv_format.h"
#include "nccl_dyn.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstring>
#include <string>
#include <vector>

using namespace bc;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Staging {                 // device copies of one host batch
    DevBuf ref_read_off, starts, cigar_off, cigar, seq_woff, planes, okmask, exc_read, exc_pos, chunks;
    DevBuf deferred;             // chunks k1_count_fast leaves to the general walker (device-written)
    DevBuf ctrl;                 // {deferred chunks, walker warps done, chunks of the last walker launch, -}: see k1_count_tiled
    Chunk *h_chunks = nullptr;   // pinned
    size_t h_chunks_cap = 0;
    cudaEvent_t copied = nullptr, done = nullptr, checked = nullptr;   // checked: the overflow check (side stream) is over
    bool used = false;
};

struct Resident {                // a batch kept in HBM by bc_batch_upload
    Staging st;
    BatchView view;
    uint32_t n_chunks = 0;
    int G = 32;
    uint32_t mean_words = 1;
    bool live = false;
    // What k1_count_fast defers is a function of the batch alone, so a resident batch that deferred nothing the first
    // time never will: its later launches leave out the (then empty) walker launch behind the fast kernel.
    int walker = 0;              // 0 = not known yet, 1 = needed, 2 = not needed
    bool launched = false;       // launched since the last bc_sync while walker == 0 or 2
};

}  // namespace

struct bc_handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t copy = nullptr, compute = nullptr, side = nullptr, stats = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, counted = nullptr, checked = nullptr;
    // The asynchronous --summarise reduction (bc_summary_async) runs on its own low-priority stream behind everything
    // queued on the compute stream so far, so that in a pipeline of steps the summary of step i runs beside the
    // counting kernel of step i + 1 (which, after bc_reset, writes the OTHER set of accumulators) instead of between
    // two counting kernels.  stats_busy: work is queued there that the compute stream has not waited for; stats_set:
    // the accumulator set it reads.  Anything that writes that set, a slot length or the result arena joins first.
    cudaEvent_t stats_ready = nullptr, stats_done = nullptr;
    bool stats_busy = false;
    bool stats_overlap = true;            // BASECOUNT_B200_SUMMARY_STREAM=0: summaries stay on the compute stream (A/B runs)
    bool join_before_count = false;       // BASECOUNT_B200_JOIN_K1=1: every counting kernel waits for the queued summaries (A/B runs)
    int stats_set = 0;
    cudaEvent_t set_free_stats[2] = {nullptr, nullptr};
    bool set_free_stats_valid[2] = {false, false};
    std::string err;

    uint32_t n_refs = 0;
    std::vector<uint32_t> ref_len, col_base, slot_cap;      // slot_cap: a slot's length at bc_begin (bc_truncate may not exceed it)
    uint64_t stride = 0;
    // Two sets of accumulators: bc_reset switches to the other one, which is zeroed on the side stream while the
    // kernels that still read the first (the summary of the step before) run, so neither the memset nor the sparse
    // corrections of the next batch (they need zeroed planes) sit between one step's summary and the next one's K1.
    uint32_t *d_counts = nullptr;            // = d_counts_set[cur_set]
    uint32_t *d_counts_set[2] = {nullptr, nullptr};
    int cur_set = 0;
    cudaEvent_t set_free[2] = {nullptr, nullptr}, set_zeroed[2] = {nullptr, nullptr};
    unsigned long long *d_counts64 = nullptr;
    uint32_t *d_col_base = nullptr, *d_ref_len = nullptr;
    uint32_t *d_status = nullptr;
    uint32_t *h_status = nullptr;            // pinned
    uint64_t reads_since_fold = 0;

    Staging stage[2];
    uint64_t pushes = 0;
    std::vector<Resident *> resident;

    DevBuf scratch_cov, scratch_pc, scratch_ent, scratch_sec, scratch_flags, scratch_i64, scratch_misc;
    // the amplicon windows of the last bc_amplicons call stay on the device: a pipeline of steps over one scheme
    // (--summarise-with-bed on sample after sample) uploads them once
    DevBuf tiles_dev;
    std::vector<int32_t> tiles_host;
    SummaryPartial *d_partials = nullptr;
    size_t partials_cap = 0;
    // asynchronous summaries: k2_summary writes its scalars into a device arena (a write to mapped host
    // memory kept every launch waiting ~15 us for the PCIe round trip); the arena comes back in ONE copy at
    // the next synchronisation and the values are handed to the callers' arrays there
    struct PendingCopy { size_t off; size_t bytes; void *dst; };   // arena bytes owed to a caller's array
    std::vector<PendingCopy> pending;
    // summarise partials owed to a caller as per-slot sums: slot r's partials are `count[r]` SummaryPartial records at
    // arena offset off + first[r] * sizeof(SummaryPartial), added on the host in index order
    struct PendingReduce { size_t off; std::vector<uint32_t> first, count; int64_t *nz, *cs; double *es; };
    std::vector<PendingReduce> pending_reduce;
    std::vector<uint32_t> part_off_host;  // first partial of every slot (host copy of d_part_off)
    size_t total_partials = 0;
    char *d_results = nullptr, *h_results = nullptr;         // device arena and its pinned mirror
    size_t results_cap = 0, results_used = 0;
    double *d_log2_tab = nullptr;         // log2 of small integers for the summarise reductions (k2_stats.cuh)
    uint32_t *d_part_off = nullptr;       // first partial of every slot (see summary_blocks)
    uint32_t part_off_refs = 0;           // 0 = stale (slot lengths changed)
    uint32_t part_off_cap = 0;            // slots the offset / arrival arrays were allocated for
    uint32_t summary_max_blocks = 1;

    cudaEvent_t t0 = nullptr, t1 = nullptr;
    static constexpr int kHist = 256;           // ring of (start, stop) events around the counting kernel
    cudaEvent_t k0[kHist] = {}, k1[kHist] = {};
    uint64_t k_count = 0;
    uint64_t launches = 0;
    int variant = 0;             // 0 = k1_count_fast + general walker for what it defers, 1 = per-base atomics, 2 = walker only
    int default_variant = 0;     // what variant 0 means (BASECOUNT_B200_K1=walker makes it 2)
    int walker_ctas_per_sm = 1;           // grid of the walker behind k1_count_fast: 1 until a batch deferred a lot
    int walker_max_ctas = 1;              // its occupancy limit
    uint64_t reads_since_sync = 0;
    // region sharding across GPUs: an NCCL communicator owned by the handle (bc_comm_init); halo columns and the
    // summarise scalars travel on the compute stream, no host synchronisation in between
    ncclComm_t comm = nullptr;
    int comm_world = 1, comm_rank = 0;
    DevBuf halo_recv, comm_scratch;
    bool side_needs_compute = true;       // the next corrections kernel must wait for everything queued on the compute stream
    int dbg_skip = 0;                     // BASECOUNT_B200_DEBUG_SKIP (timing experiments only): 1 = no walker launch, 2 = no corrections
};

#define CU(h, expr)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
            return BC_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

// The compute stream waits for what is queued on the summary stream (see bc_handle::stats).
static int join_stats(bc_handle *h)
{
    if (h->stats_busy) {
        CU(h, cudaStreamWaitEvent(h->compute, h->stats_done, 0));
        h->stats_busy = false;
    }
    return BC_OK;
}
#define JOIN_STATS(h)                 \
    do {                              \
        int rcj__ = join_stats(h);    \
        if (rcj__) return rcj__;      \
    } while (0)

static int fail(bc_handle *h, int code, const char *msg)
{
    h->err = msg;
    return code;
}

static int ensure(bc_handle *h, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return BC_OK;
    if (b.p) {
        CU(h, cudaDeviceSynchronize());
        CU(h, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    CU(h, cudaMalloc(&b.p, want));
    b.cap = want;
    return BC_OK;
}

// Queue the copy of the result arena to its pinned mirror (no-op without pending summaries).
static int fetch_summaries(bc_handle *h)
{
    if (h->pending.empty() && h->pending_reduce.empty()) return BC_OK;
    JOIN_STATS(h);
    CU(h, cudaMemcpyAsync(h->h_results, h->d_results, h->results_used, cudaMemcpyDeviceToHost, h->compute));
    return BC_OK;
}

// Call after fetch_summaries and a synchronisation of the compute stream: hands the values to their callers.
static void deliver_summaries(bc_handle *h)
{
    for (auto &p : h->pending) std::memcpy(p.dst, h->h_results + p.off, p.bytes);
    h->pending.clear();
    for (auto &q : h->pending_reduce) {
        const SummaryPartial *all = reinterpret_cast<const SummaryPartial *>(h->h_results + q.off);
        for (size_t r = 0; r < q.first.size(); r++) {
            long long nz = 0, cs = 0;
            double es = 0.0;
            for (uint32_t i = 0; i < q.count[r]; i++) {
                const SummaryPartial &sp = all[q.first[r] + i];
                nz += sp.nonzero;
                cs += sp.cov_sum;
                es += sp.ent_sum;
            }
            q.nz[r] = nz;
            q.cs[r] = cs;
            q.es[r] = es;
        }
    }
    h->pending_reduce.clear();
    h->results_used = 0;
}

// Room for `bytes` more in the arena; when it is full the pending summaries are delivered early.
static int reserve_results(bc_handle *h, size_t bytes, size_t *off)
{
    if (h->results_used + bytes > h->results_cap) {
        int rc = fetch_summaries(h);
        if (rc) return rc;
        CU(h, cudaStreamSynchronize(h->compute));
        deliver_summaries(h);
        if (bytes > h->results_cap) {
            if (h->d_results) CU(h, cudaFree(h->d_results));
            if (h->h_results) CU(h, cudaFreeHost(h->h_results));
            h->d_results = h->h_results = nullptr;
            h->results_cap = 0;
            const size_t want = std::max<size_t>(bytes * 4, 1u << 20);
            CU(h, cudaMalloc((void **)&h->d_results, want));
            CU(h, cudaHostAlloc((void **)&h->h_results, want, cudaHostAllocDefault));
            h->results_cap = want;
        }
    }
    *off = h->results_used;
    h->results_used += bytes;
    return BC_OK;
}

static void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static void release_staging(Staging &s)
{
    release(s.ref_read_off); release(s.starts); release(s.cigar_off); release(s.cigar); release(s.seq_woff);
    release(s.planes); release(s.okmask); release(s.exc_read); release(s.exc_pos); release(s.chunks); release(s.deferred); release(s.ctrl);
    if (s.h_chunks) cudaFreeHost(s.h_chunks);
    s.h_chunks = nullptr;
    s.h_chunks_cap = 0;
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.done) cudaEventDestroy(s.done);
    if (s.checked) cudaEventDestroy(s.checked);
    s.copied = s.done = s.checked = nullptr;
}

extern "C" {

int bc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bc_device_pci_bus_id(int device, char *out, int len)
{
    if (!out || len < 16) return BC_ERR_ARG;
    if (cudaDeviceGetPCIBusId(out, len, device) != cudaSuccess) return BC_ERR_CUDA;
    for (char *p = out; *p; p++)
        if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    return BC_OK;
}

const char *bc_last_error(bc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int bc_create(int device, bc_handle **out)
{
    if (!out) return BC_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (basecount_b200 has no CPU path)";
        return BC_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        g_create_error = "device index out of range";
        return BC_ERR_ARG;
    }
    bc_handle *h = new bc_handle();
    h->device = device;
    auto bail = [&](cudaError_t ce, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete h;
        return BC_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    h->sm_count = prop.multiProcessorCount;
    if (const char *k1 = std::getenv("BASECOUNT_B200_K1")) {       // experiments: which tiled kernel "variant 0" runs
        if (std::strcmp(k1, "walker") == 0) h->default_variant = 2;
        else if (std::strcmp(k1, "fast") != 0) return bail(cudaErrorInvalidValue, "BASECOUNT_B200_K1 must be fast or walker");
        h->variant = h->default_variant;
    }
    if (const char *d = std::getenv("BASECOUNT_B200_DEBUG_SKIP")) h->dbg_skip = std::atoi(d);
    if ((e = cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    {
        // K1 is one resident wave whose three CTAs per SM hold 96 % of the register file: whatever shares
        // the SMs with it takes the place of K1 CTAs, which then start a wave late.  The side stream
        // (sparse corrections, overflow check) therefore has the LOWEST priority and its kernels are
        // launched AFTER K1: their CTAs go where K1's CTAs leave room (measured: step 123.2 -> 119.4 us,
        // K1 93.8 -> 90.3 us; the launch order alone changes nothing).
        int least = 0, greatest = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&least, &greatest)) != cudaSuccess) return bail(e, "priority range");
        if ((e = cudaStreamCreateWithPriority(&h->compute, cudaStreamNonBlocking, greatest)) != cudaSuccess)
            return bail(e, "stream");
        if ((e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, least)) != cudaSuccess)
            return bail(e, "stream");
        if ((e = cudaStreamCreateWithPriority(&h->stats, cudaStreamNonBlocking, least)) != cudaSuccess)
            return bail(e, "stream");
    }
    if (const char *o = std::getenv("BASECOUNT_B200_SUMMARY_STREAM")) h->stats_overlap = std::atoi(o) != 0;
    if (const char *o = std::getenv("BASECOUNT_B200_JOIN_K1")) h->join_before_count = std::atoi(o) != 0;
    if ((e = cudaEventCreateWithFlags(&h->stats_ready, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->stats_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->counted, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    if ((e = cudaEventCreateWithFlags(&h->checked, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    for (int i = 0; i < 2; i++) {
        if ((e = cudaEventCreateWithFlags(&h->set_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreateWithFlags(&h->set_zeroed[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreateWithFlags(&h->set_free_stats[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    }
    for (cudaEvent_t *ev : {&h->t0, &h->t1})
        if ((e = cudaEventCreate(ev)) != cudaSuccess) return bail(e, "event");
    for (int i = 0; i < bc_handle::kHist; i++) {
        if ((e = cudaEventCreate(&h->k0[i])) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreate(&h->k1[i])) != cudaSuccess) return bail(e, "event");
    }
    for (Staging &s : h->stage) {
        if ((e = cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
        if ((e = cudaEventCreateWithFlags(&s.checked, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    }
    if ((e = cudaMalloc(&h->d_status, kStatWords * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemset(h->d_status, 0, kStatWords * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMemset");
    if ((e = cudaMalloc(&h->d_log2_tab, kSummaryTabDoubles * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc");
    k_fill_log2<<<(kSummaryTabDoubles + 255) / 256, 256, 0, h->compute>>>(h->d_log2_tab);
    if ((e = cudaStreamSynchronize(h->compute)) != cudaSuccess) return bail(e, "k_fill_log2");
    if ((e = cudaHostAlloc((void **)&h->h_status, kStatWords * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess)
        return bail(e, "cudaHostAlloc");
    *out = h;
    return BC_OK;
}

void bc_destroy(bc_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (Staging &s : h->stage) release_staging(s);
    for (Resident *r : h->resident) {
        if (r) {
            release_staging(r->st);
            delete r;
        }
    }
    for (DevBuf *b : {&h->scratch_cov, &h->scratch_pc, &h->scratch_ent, &h->scratch_sec, &h->scratch_flags,
                      &h->scratch_i64, &h->scratch_misc, &h->tiles_dev})
        release(*b);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_part_off) cudaFree(h->d_part_off);
    if (h->d_results) cudaFree(h->d_results);
    if (h->h_results) cudaFreeHost(h->h_results);
    for (int i = 0; i < 2; i++) {
        if (h->d_counts_set[i]) cudaFree(h->d_counts_set[i]);
        if (h->set_free[i]) cudaEventDestroy(h->set_free[i]);
        if (h->set_zeroed[i]) cudaEventDestroy(h->set_zeroed[i]);
        if (h->set_free_stats[i]) cudaEventDestroy(h->set_free_stats[i]);
    }
    if (h->stats_ready) cudaEventDestroy(h->stats_ready);
    if (h->stats_done) cudaEventDestroy(h->stats_done);
    if (h->d_counts64) cudaFree(h->d_counts64);
    if (h->d_col_base) cudaFree(h->d_col_base);
    if (h->d_ref_len) cudaFree(h->d_ref_len);
    if (h->comm) bcnccl::api().CommDestroy(h->comm);
    release(h->halo_recv);
    release(h->comm_scratch);
    if (h->d_status) cudaFree(h->d_status);
    if (h->d_log2_tab) cudaFree(h->d_log2_tab);
    if (h->h_status) cudaFreeHost(h->h_status);
    for (cudaEvent_t ev : {h->t0, h->t1})
        if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < bc_handle::kHist; i++) {
        if (h->k0[i]) cudaEventDestroy(h->k0[i]);
        if (h->k1[i]) cudaEventDestroy(h->k1[i]);
    }
    if (h->fork) cudaEventDestroy(h->fork);
    if (h->join) cudaEventDestroy(h->join);
    if (h->counted) cudaEventDestroy(h->counted);
    if (h->checked) cudaEventDestroy(h->checked);
    if (h->copy) cudaStreamDestroy(h->copy);
    if (h->compute) cudaStreamDestroy(h->compute);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->stats) cudaStreamDestroy(h->stats);
    delete h;
}

int bc_host_alloc(size_t bytes, void **out)
{
    if (!out) return BC_ERR_ARG;
    *out = nullptr;
    if (bytes == 0) bytes = 1;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? BC_OK : BC_ERR_CUDA;
}

void bc_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int bc_begin(bc_handle *h, uint32_t n_refs, const uint32_t *ref_lens)
{
    if (!h) return BC_ERR_ARG;
    if (n_refs == 0 || !ref_lens) return fail(h, BC_ERR_ARG, "bc_begin: need at least one reference");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaDeviceSynchronize());
    h->stats_busy = false;
    h->set_free_stats_valid[0] = h->set_free_stats_valid[1] = false;
    uint64_t total = 0;
    std::vector<uint32_t> cb(n_refs), rl(n_refs);
    for (uint32_t r = 0; r < n_refs; r++) {
        if (ref_lens[r] > 0x7FFFFFFFu) return fail(h, BC_ERR_ARG, "bc_begin: reference longer than 2^31-1 (BAM l_ref limit)");
        rl[r] = ref_lens[r];
        if (total > 0xFFFFFFFFull - kColAlign) return fail(h, BC_ERR_ARG, "bc_begin: more than 2^32 columns in one handle");
        cb[r] = (uint32_t)total;
        total += ((uint64_t)ref_lens[r] + kColAlign - 1) / kColAlign * kColAlign;
    }
    total += 32u * kW * 32u + 64u;   // slack: a full window (64 * G columns, G <= 32) may overhang the last slot
    if (total > 0xFFFFFFFFull) return fail(h, BC_ERR_ARG, "bc_begin: more than 2^32 columns in one handle");
    if (total != h->stride || n_refs != h->n_refs) {
        for (int i = 0; i < 2; i++) {
            if (h->d_counts_set[i]) CU(h, cudaFree(h->d_counts_set[i]));
            h->d_counts_set[i] = nullptr;
        }
        if (h->d_counts64) CU(h, cudaFree(h->d_counts64));
        if (h->d_col_base) CU(h, cudaFree(h->d_col_base));
        if (h->d_ref_len) CU(h, cudaFree(h->d_ref_len));
        h->d_counts = nullptr;
        h->d_counts64 = nullptr;
        h->d_col_base = h->d_ref_len = nullptr;
        for (int i = 0; i < 2; i++) CU(h, cudaMalloc(&h->d_counts_set[i], (size_t)total * kPlanes * sizeof(uint32_t)));
        CU(h, cudaMalloc(&h->d_col_base, n_refs * sizeof(uint32_t)));
        CU(h, cudaMalloc(&h->d_ref_len, n_refs * sizeof(uint32_t)));
    } else if (h->d_counts64) {
        CU(h, cudaFree(h->d_counts64));
        h->d_counts64 = nullptr;
    }
    h->stride = total;
    h->n_refs = n_refs;
    h->part_off_refs = 0;
    h->ref_len = rl;
    h->slot_cap = rl;
    h->col_base = cb;
    h->reads_since_fold = 0;
    h->cur_set = 0;
    h->side_needs_compute = true;
    h->d_counts = h->d_counts_set[0];
    for (int i = 0; i < 2; i++)
        CU(h, cudaMemsetAsync(h->d_counts_set[i], 0, (size_t)total * kPlanes * sizeof(uint32_t), h->compute));
    CU(h, cudaMemsetAsync(h->d_status, 0, kStatWords * sizeof(uint32_t), h->compute));
    CU(h, cudaMemcpyAsync(h->d_col_base, cb.data(), n_refs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->compute));
    CU(h, cudaMemcpyAsync(h->d_ref_len, rl.data(), n_refs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

int bc_reset(bc_handle *h)
{
    if (!h) return BC_ERR_ARG;
    if (h->n_refs == 0) return fail(h, BC_ERR_STATE, "bc_begin has not been called");
    CU(h, cudaSetDevice(h->device));
    // everything queued so far may still read the current set; the other one is free once what was queued before
    // the previous switch is done.  It is zeroed on the side stream, behind the overflow check of the last batch
    // (i.e. not beside its K1), and the compute stream only waits for that memset.
    CU(h, cudaEventRecord(h->set_free[h->cur_set], h->compute));
    // ... and so may the summaries queued on their own stream (all of them read the set in use when they were queued)
    h->set_free_stats_valid[h->cur_set] = h->stats_busy;
    if (h->stats_busy) CU(h, cudaEventRecord(h->set_free_stats[h->cur_set], h->stats));
    h->cur_set ^= 1;
    h->d_counts = h->d_counts_set[h->cur_set];
    CU(h, cudaStreamWaitEvent(h->side, h->set_free[h->cur_set], 0));
    if (h->set_free_stats_valid[h->cur_set]) {
        CU(h, cudaStreamWaitEvent(h->side, h->set_free_stats[h->cur_set], 0));
        h->set_free_stats_valid[h->cur_set] = false;
    }
    CU(h, cudaMemsetAsync(h->d_counts, 0, (size_t)h->stride * kPlanes * sizeof(uint32_t), h->side));
    CU(h, cudaEventRecord(h->set_zeroed[h->cur_set], h->side));
    CU(h, cudaStreamWaitEvent(h->compute, h->set_zeroed[h->cur_set], 0));
    if (h->d_counts64) {                                 // (one copy for both sets: a summary may still read it)
        JOIN_STATS(h);
        CU(h, cudaMemsetAsync(h->d_counts64, 0, (size_t)h->stride * kPlanes * sizeof(unsigned long long), h->compute));
    }
    h->reads_since_fold = 0;
    return BC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ batches
// Lanes per read slot.  The window (64 * G columns) must hold a typical read at any 32-column
// alignment; beyond that a wider window means fewer flushes (one per window of reads) but fewer
// pieces per trip, so pick the G that minimises an instruction estimate per read:
//     flush(G) / reads per flush  +  trip / pieces per trip
// with reads per flush = what starts inside one window at the batch's mean depth, capped by the
// 252-pieces-per-slot counter limit.  Deep amplicon piles end up at the narrowest G that fits the
// reads; shallow whole-genome coverage (30x) at the next one up.
static int pick_group_width(const bc_handle *h, const bc_batch *b, uint64_t total_words)
{
    if (const char *g = std::getenv("BASECOUNT_B200_G")) {         // experiments: force the lane-group width
        const int v = std::atoi(g);
        if (v == 4 || v == 8 || v == 16 || v == 32) return v;
    }
    uint64_t mean = b->mean_read_len;
    if (mean == 0 && b->n_reads) mean = total_words * 32u / b->n_reads;
    if (mean == 0) mean = 1;
    uint64_t cols = 0;                                           // reference columns of the slots that got reads
    for (uint32_t r = 0; r < b->n_refs; r++)
        if (b->ref_read_off[r + 1] > b->ref_read_off[r]) cols += h->ref_len[r];
    const double depth = cols ? (double)b->n_reads * (double)mean / (double)cols : 1e9;
    static const double flush_cost[4] = {1000.0, 1500.0, 2300.0, 4000.0};      // G = 4, 8, 16, 32 (profiled / estimated)
    int best = 32;
    double best_cost = 1e300;
    for (int i = 0; i < 4; i++) {
        const int G = 4 << i;
        const double win = 32.0 * kW * G;
        if (G < 32 && (double)(mean + mean / 8) > win - 31.0) continue;          // reads would not fit
        const int S = 32 / G;
        const double fit = std::max(win - 31.0 - (double)mean, 32.0);           // start positions that fit one window
        const double per_window = std::max(1.0, depth * fit / (double)mean);
        const double per_flush = std::min(per_window, 252.0 * S / 1.1);
        const double cost = flush_cost[i] / per_flush + 240.0 / (4.0 * S);
        if (cost < best_cost) {
            best_cost = cost;
            best = G;
        }
    }
    return best;
}

template <int G, bool OK>
static int k1_prepare(size_t *smem, int *ctas_per_sm)
{
    *smem = (size_t)k1_cta_smem_bytes<G, OK>();
    cudaError_t e = cudaFuncSetAttribute(k1_count_tiled<G, OK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k1_count_tiled<G, OK>, kK1Threads, *smem);
}

template <int G, bool OK>
static int k1_fast_prepare(size_t *smem, int *ctas_per_sm)
{
    *smem = (size_t)k1_fast_cta_smem_bytes<G, OK>();
    cudaError_t e = cudaFuncSetAttribute(k1_count_fast<G, OK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k1_count_fast<G, OK>, kK1Threads, *smem);
}

// Warps of the chosen K1 instance that are resident on the whole GPU at once (variant 0 also prepares the walker
// that runs behind k1_count_fast and reports its occupancy in *walker_ctas).
static int k1_resident_warps(bc_handle *h, int G, bool ok, uint32_t *out, int *walker_ctas = nullptr)
{
    size_t smem = 0;
    int ctas = 0, wctas = 0, e = 0;
    const bool fast = h->variant == 0;
#define K1_PREP(GG)                                                                                             \
    do {                                                                                                        \
        e = ok ? k1_prepare<GG, true>(&smem, &wctas) : k1_prepare<GG, false>(&smem, &wctas);                    \
        ctas = wctas;                                                                                           \
        if (e == 0 && fast) e = ok ? k1_fast_prepare<GG, true>(&smem, &ctas) : k1_fast_prepare<GG, false>(&smem, &ctas); \
    } while (0)
    if (G == 4) K1_PREP(4);
    else if (G == 8) K1_PREP(8);
    else if (G == 16) K1_PREP(16);
    else K1_PREP(32);
#undef K1_PREP
    if (e != 0) {
        h->err = std::string("k1 occupancy query: ") + cudaGetErrorString((cudaError_t)e);
        return BC_ERR_CUDA;
    }
    if (walker_ctas) *walker_ctas = std::max(1, wctas);
    *out = (uint32_t)std::max(1, ctas) * kK1WarpsPerCta * (uint32_t)h->sm_count;
    return BC_OK;
}

// Split every slot's reads into per-warp chunks: about one chunk per resident warp (a single
// full wave), because every chunk pays one counter flush at its end.
static uint32_t build_chunks(bc_handle *h, const uint32_t *ref_read_off, uint32_t n_reads, uint32_t target_warps,
                             std::vector<Chunk> &out)
{
    out.clear();
    uint32_t per = (uint32_t)std::max<uint64_t>(64, ((uint64_t)n_reads + target_warps - 1) / target_warps);
    for (uint32_t r = 0; r < h->n_refs; r++) {
        const uint32_t n_r = ref_read_off[r + 1] - ref_read_off[r];
        if (n_r == 0) continue;
        const uint32_t pieces = (n_r + per - 1) / per;
        const uint32_t each = (n_r + pieces - 1) / pieces;          // equal shares within the slot
        for (uint32_t a = ref_read_off[r]; a < ref_read_off[r + 1]; a += each) {
            Chunk c;
            c.read_begin = a;
            c.read_end = std::min(a + each, ref_read_off[r + 1]);
            c.col_base = h->col_base[r];
            c.ref_len = h->ref_len[r];
            out.push_back(c);
        }
    }
    return (uint32_t)out.size();
}

static int validate_batch(bc_handle *h, const bc_batch *b)
{
    if (!b) return fail(h, BC_ERR_ARG, "null batch");
    if (h->n_refs == 0) return fail(h, BC_ERR_STATE, "bc_begin has not been called");
    if (b->n_refs != h->n_refs) return fail(h, BC_ERR_ARG, "batch n_refs differs from bc_begin");
    if (b->n_reads && (!b->ref_read_off || !b->starts || !b->cigar_off || !b->seq_woff))
        return fail(h, BC_ERR_ARG, "batch is missing a required array");
    if (b->n_exc && (!b->exc_read || !b->exc_pos)) return fail(h, BC_ERR_ARG, "batch is missing exception arrays");
    return BC_OK;
}

// Launch K1 (+ corrections, + exact overflow check) for a batch whose arrays are in HBM.
static uint32_t fast_reads_per_block(uint32_t mean_words)
{
    return std::max<uint32_t>(1, std::min<uint32_t>(kFastRpbMax, (kSeqCap - 8) / std::max<uint32_t>(1, mean_words)));
}

// Stages of k1_count_fast's sequence buffer: four short ones when a block of reads fits a quarter of it.
static uint32_t fast_stages(uint32_t mean_words)
{
    const uint32_t block_words = fast_reads_per_block(mean_words) * (std::max<uint32_t>(1, mean_words) + 1) + 8;
    return block_words <= kFastStageShort ? 4u : 3u;
}

static int launch_count(bc_handle *h, const BatchView &v, const Chunk *d_chunks, uint32_t n_chunks, int G,
                        uint32_t mean_words, Chunk *d_deferred, uint32_t *d_ctrl, Resident *res, cudaEvent_t copied)
{
    if (v.n_reads == 0) return BC_OK;
    // a summary that still reads THIS set of accumulators (no bc_reset since it was queued) comes first
    if (h->stats_busy && (h->stats_set == h->cur_set || h->join_before_count)) JOIN_STATS(h);
    // uint32 counters cannot wrap while fewer than 2^32 reads went in since the last fold
    if (h->reads_since_fold + v.n_reads > 0xFFFFFFFFull) {
        const uint64_t n = h->stride * kPlanes;
        JOIN_STATS(h);
        if (!h->d_counts64) {
            CU(h, cudaMalloc(&h->d_counts64, n * sizeof(unsigned long long)));
            CU(h, cudaMemsetAsync(h->d_counts64, 0, n * sizeof(unsigned long long), h->compute));
        }
        k_fold_counts<<<(unsigned)((n + 255) / 256), 256, 0, h->compute>>>(h->d_counts, h->d_counts64, n);
        h->launches++;
        h->reads_since_fold = 0;
        h->side_needs_compute = true;
    }
    h->reads_since_fold += v.n_reads;
    h->reads_since_sync += v.n_reads;

    CountView cv;
    cv.counts = h->d_counts;
    cv.stride = h->stride;
    cv.col_base = h->d_col_base;
    cv.ref_len = h->d_ref_len;
    cv.status = h->d_status;

    // The sparse corrections only add to planes A and N with atomics, so they commute with K1 and need nothing from
    // the compute stream but zeroed planes (bc_reset zeroes on the side stream, i.e. in front of them) and the batch
    // itself.  They go to the side (lowest-priority) stream and wait there for the K1 of the batch BEFORE this one,
    // so in a pipeline of steps they run beside that step's summary kernel, not beside a counting kernel (beside K1
    // they cost it ~5 us).  After an operation that does not commute (truncate, halo add, fold, begin) they wait
    // for the whole compute stream once.  The exact overflow check follows on the same stream once this K1 is done:
    // it only writes status words (read in bc_sync), so the statistics kernels need not wait for it.
    if (v.n_exc && !(h->dbg_skip & 2)) {
        if (h->side_needs_compute) {
            CU(h, cudaEventRecord(h->fork, h->compute));
            CU(h, cudaStreamWaitEvent(h->side, h->fork, 0));
            h->side_needs_compute = false;
        } else {
            if (copied) CU(h, cudaStreamWaitEvent(h->side, copied, 0));
            CU(h, cudaStreamWaitEvent(h->side, h->counted, 0));          // the previous K1 (no-op if there was none)
        }
        k1_exceptions<<<(v.n_exc + 127) / 128, 128, 0, h->side>>>(v, cv);
        CU(h, cudaEventRecord(h->join, h->side));
        h->launches++;
    }
    const int ki = (int)(h->k_count % bc_handle::kHist);
    CU(h, cudaEventRecord(h->k0[ki], h->compute));
    if (h->variant == 1) {
        k1_count_per_base<<<(v.n_reads + 127) / 128, 128, 0, h->compute>>>(v, cv);
    } else {
        const unsigned grid = (n_chunks + kK1WarpsPerCta - 1) / kK1WarpsPerCta;
        const bool ok = v.okmask != nullptr;
        // reads per staging block: a block's plane words must fit one pipeline stage
        const uint32_t words_per_read = std::max<uint32_t>(1, mean_words);
        const uint32_t rpb = std::max<uint32_t>(1, std::min<uint32_t>(kMaxRpb, (kSeqCap - 8) / words_per_read));
        const uint32_t rpb_fast = fast_reads_per_block(mean_words);
        const bool fast = h->variant == 0;
        // variant 0: k1_count_fast takes every block of reads its straight-line decode can and defers the rest, as
        // chunks, to a list in HBM; the general walker runs right behind it over that list (a fixed grid striding
        // over a device-side count: with nothing deferred it is an empty launch).
        const unsigned wgrid = (unsigned)(h->sm_count * std::max(1, h->walker_ctas_per_sm));
        const bool walk = !(res && res->walker == 2) && !(h->dbg_skip & 1);
        if (res && res->walker != 1) res->launched = true;
#define K1_LAUNCH(GG, OKK)                                                                                              \
    do {                                                                                                                \
        const size_t wsmem = (size_t)k1_cta_smem_bytes<GG, OKK>();                                                      \
        if (fast) {                                                                                                     \
            const size_t smem = (size_t)k1_fast_cta_smem_bytes<GG, OKK>();                                              \
            k1_count_fast<GG, OKK><<<grid, kK1Threads, smem, h->compute>>>(v, cv, d_chunks, n_chunks, rpb_fast,         \
                                                                            fast_stages(mean_words), d_deferred,       \
                                                                            d_ctrl);                                    \
            if (walk) {                                                                                                 \
                k1_count_tiled<GG, OKK><<<wgrid, kK1Threads, wsmem, h->compute>>>(v, cv, d_deferred, 0u, rpb, d_ctrl);   \
                h->launches++;                                                                                          \
            }                                                                                                           \
        } else {                                                                                                        \
            k1_count_tiled<GG, OKK><<<grid, kK1Threads, wsmem, h->compute>>>(v, cv, d_chunks, n_chunks, rpb, nullptr);   \
        }                                                                                                               \
    } while (0)
#define K1_LAUNCH_G(GG)                 \
    do {                                \
        if (ok) K1_LAUNCH(GG, true);    \
        else K1_LAUNCH(GG, false);      \
    } while (0)
        if (G == 4) K1_LAUNCH_G(4);
        else if (G == 8) K1_LAUNCH_G(8);
        else if (G == 16) K1_LAUNCH_G(16);
        else K1_LAUNCH_G(32);
#undef K1_LAUNCH_G
#undef K1_LAUNCH
    }
    CU(h, cudaEventRecord(h->k1[ki], h->compute));
    h->k_count++;
    h->launches++;
    CU(h, cudaEventRecord(h->counted, h->compute));
    CU(h, cudaStreamWaitEvent(h->side, h->counted, 0));
    // (128-thread CTAs: 3,328 registers, small enough to run on an SM beside the three CTAs of the next counting kernel)
    k1_check_overflow<<<std::min<unsigned>((v.n_reads + 127) / 128, (unsigned)h->sm_count * 4), 128, 0, h->side>>>(v, cv);
    CU(h, cudaEventRecord(h->checked, h->side));
    h->launches++;
    if (v.n_exc && !(h->dbg_skip & 2)) CU(h, cudaStreamWaitEvent(h->compute, h->join, 0));
    CU(h, cudaGetLastError());
    return BC_OK;
}

static int stage_copy(bc_handle *h, DevBuf &dst, const void *src, size_t bytes, bool src_on_device)
{
    if (bytes == 0) return BC_OK;
    int rc = ensure(h, dst, bytes);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(dst.p, src, bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->copy));
    return BC_OK;
}

// Copy a host batch into `st` on the copy stream and fill `view` / chunk table.
static int stage_batch(bc_handle *h, Staging &st, const bc_batch *b, BatchView &view, uint32_t &n_chunks, int &G,
                       uint32_t &mean_words)
{
    const uint32_t n = b->n_reads;
    view = BatchView();
    view.n_reads = n;
    view.n_refs = b->n_refs;
    view.n_exc = b->n_exc;
    n_chunks = 0;
    G = 32;
    mean_words = 1;
    if (n == 0) return BC_OK;
    if (n > 0xFFFFFFFFu - 256u) return fail(h, BC_ERR_ARG, "too many reads in one batch");
    if (b->ref_read_off[0] != 0 || b->ref_read_off[b->n_refs] != n)
        return fail(h, BC_ERR_ARG, "ref_read_off must start at 0 and end at n_reads");
    for (uint32_t r = 0; r < b->n_refs; r++)
        if (b->ref_read_off[r] > b->ref_read_off[r + 1]) return fail(h, BC_ERR_ARG, "ref_read_off must be non-decreasing");
    const uint64_t n_cigar = b->cigar_off[n];
    const uint64_t n_words = b->seq_woff[n];
    if (n_cigar && !b->cigar) return fail(h, BC_ERR_ARG, "batch is missing cigar words");
    if (n_words && !b->planes) return fail(h, BC_ERR_ARG, "batch is missing sequence planes");
    G = pick_group_width(h, b, n_words);
    mean_words = (uint32_t)((n_words + n - 1) / n);

    std::vector<Chunk> chunks;
    uint32_t target_warps = 0;
    int rc0 = k1_resident_warps(h, G, b->okmask != nullptr, &target_warps, &h->walker_max_ctas);
    if (rc0) return rc0;
    n_chunks = build_chunks(h, b->ref_read_off, n, target_warps, chunks);
    // room for every run of blocks k1_count_fast may defer (at worst every other block of every chunk)
    if ((rc0 = ensure(h, st.deferred, ((size_t)n / fast_reads_per_block(mean_words) + n_chunks + 16) * sizeof(Chunk)))) return rc0;
    if (!st.ctrl.p) {
        if ((rc0 = ensure(h, st.ctrl, 4 * sizeof(uint32_t)))) return rc0;
        CU(h, cudaMemsetAsync(st.ctrl.p, 0, 4 * sizeof(uint32_t), h->copy));
    }
    if (n_chunks > st.h_chunks_cap) {
        if (st.h_chunks) CU(h, cudaFreeHost(st.h_chunks));
        st.h_chunks = nullptr;
        st.h_chunks_cap = 0;
        size_t want = (size_t)n_chunks + n_chunks / 2 + 64;
        CU(h, cudaHostAlloc((void **)&st.h_chunks, want * sizeof(Chunk), cudaHostAllocDefault));
        st.h_chunks_cap = want;
    }
    std::memcpy(st.h_chunks, chunks.data(), (size_t)n_chunks * sizeof(Chunk));

    int rc;
    if ((rc = stage_copy(h, st.ref_read_off, b->ref_read_off, (size_t)(b->n_refs + 1) * 4, false))) return rc;
    if ((rc = stage_copy(h, st.starts, b->starts, (size_t)n * 4, false))) return rc;
    if ((rc = stage_copy(h, st.cigar_off, b->cigar_off, (size_t)(n + 1) * 4, false))) return rc;
    if ((rc = stage_copy(h, st.cigar, b->cigar, (size_t)n_cigar * 4, false))) return rc;
    if ((rc = stage_copy(h, st.seq_woff, b->seq_woff, (size_t)(n + 1) * 4, false))) return rc;
    if ((rc = stage_copy(h, st.planes, b->planes, (size_t)n_words * 8, false))) return rc;
    if (b->okmask && (rc = stage_copy(h, st.okmask, b->okmask, (size_t)n_words * 4, false))) return rc;
    if (b->n_exc) {
        if ((rc = stage_copy(h, st.exc_read, b->exc_read, (size_t)b->n_exc * 4, false))) return rc;
        if ((rc = stage_copy(h, st.exc_pos, b->exc_pos, (size_t)b->n_exc * 4, false))) return rc;
    }
    if ((rc = stage_copy(h, st.chunks, st.h_chunks, (size_t)n_chunks * sizeof(Chunk), false))) return rc;

    view.ref_read_off = (const uint32_t *)st.ref_read_off.p;
    view.starts = (const uint32_t *)st.starts.p;
    view.cigar_off = (const uint32_t *)st.cigar_off.p;
    view.cigar = (const uint32_t *)st.cigar.p;
    view.seq_woff = (const uint32_t *)st.seq_woff.p;
    view.planes = (const uint2 *)st.planes.p;
    view.okmask = b->okmask ? (const uint32_t *)st.okmask.p : nullptr;
    view.exc_read = (const uint32_t *)st.exc_read.p;
    view.exc_pos = (const uint32_t *)st.exc_pos.p;
    return BC_OK;
}

extern "C" {

int bc_push_batch(bc_handle *h, const bc_batch *b)
{
    if (!h) return BC_ERR_ARG;
    int rc = validate_batch(h, b);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    if (b->on_device) {
        const uint32_t id = b->reserved;
        if (id == 0 || id > h->resident.size() || !h->resident[id - 1] || !h->resident[id - 1]->live)
            return fail(h, BC_ERR_ARG, "on_device batch was not created by bc_batch_upload on this handle");
        Resident *r = h->resident[id - 1];
        return launch_count(h, r->view, (const Chunk *)r->st.chunks.p, r->n_chunks, r->G, r->mean_words,
                            (Chunk *)r->st.deferred.p, (uint32_t *)r->st.ctrl.p, r, nullptr);
    }
    Staging &st = h->stage[h->pushes & 1];
    h->pushes++;
    if (st.used) {
        // the kernels that last read this staging set must be done before it is overwritten
        CU(h, cudaEventSynchronize(st.done));
        CU(h, cudaEventSynchronize(st.checked));
    }
    BatchView view;
    uint32_t n_chunks, mean_words;
    int G;
    if ((rc = stage_batch(h, st, b, view, n_chunks, G, mean_words))) return rc;
    CU(h, cudaEventRecord(st.copied, h->copy));
    CU(h, cudaStreamWaitEvent(h->compute, st.copied, 0));
    rc = launch_count(h, view, (const Chunk *)st.chunks.p, n_chunks, G, mean_words, (Chunk *)st.deferred.p,
                      (uint32_t *)st.ctrl.p, nullptr, st.copied);
    CU(h, cudaEventRecord(st.done, h->compute));
    CU(h, cudaEventRecord(st.checked, h->side));
    st.used = true;
    return rc;
}

int bc_sync(bc_handle *h)
{
    if (!h) return BC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->copy));
    CU(h, cudaStreamSynchronize(h->side));               // the overflow checks have written their status words
    {
        int rcf = fetch_summaries(h);
        if (rcf) return rcf;
    }
    CU(h, cudaMemcpyAsync(h->h_status, h->d_status, kStatWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    deliver_summaries(h);
    for (Resident *r : h->resident) {
        if (!r || !r->live || !r->launched) continue;
        uint32_t c[4] = {0, 0, 0, 0};
        CU(h, cudaMemcpy(c, r->st.ctrl.p, sizeof(c), cudaMemcpyDeviceToHost));
        r->launched = false;
        if (r->walker == 0) r->walker = c[2] ? 1 : 2;
        else if (r->walker == 2 && c[0])
            return fail(h, BC_ERR_STATE, "internal: a resident batch deferred reads after its walker launch was dropped");
    }
    {
        // size the walker's grid for the next batches from what this one deferred: one CTA per SM while nearly
        // everything takes the fast kernel (an empty launch then), full occupancy once an eighth of the reads do not
        const uint64_t deferred = h->h_status[kStatDeferredReads];
        h->walker_ctas_per_sm = (deferred * 8u > h->reads_since_sync) ? h->walker_max_ctas : 1;
        h->reads_since_sync = 0;
        if (deferred && !h->h_status[kStatIndexError] && !h->h_status[kStatMaybeOverflow])
            CU(h, cudaMemsetAsync(h->d_status, 0, kStatWords * sizeof(uint32_t), h->compute));
    }
    if (h->h_status[kStatIndexError]) {
        CU(h, cudaMemsetAsync(h->d_status, 0, kStatWords * sizeof(uint32_t), h->compute));
        return fail(h, BC_ERR_INDEX, "alignment counted past the end of the reference (std::out_of_range in count.cpp)");
    }
    if (h->h_status[kStatMaybeOverflow]) {
        // pieces crossed ref_len but nothing countable lay beyond it: not an error in the reference
        CU(h, cudaMemsetAsync(h->d_status, 0, kStatWords * sizeof(uint32_t), h->compute));
    }
    return BC_OK;
}

int bc_batch_upload(bc_handle *h, const bc_batch *host, bc_batch *dev)
{
    if (!h || !dev) return BC_ERR_ARG;
    int rc = validate_batch(h, host);
    if (rc) return rc;
    if (host->on_device) return fail(h, BC_ERR_ARG, "bc_batch_upload expects host pointers");
    CU(h, cudaSetDevice(h->device));
    Resident *r = new Resident();
    rc = stage_batch(h, r->st, host, r->view, r->n_chunks, r->G, r->mean_words);
    if (rc == BC_OK && cudaStreamSynchronize(h->copy) != cudaSuccess) rc = fail(h, BC_ERR_CUDA, "upload failed");
    if (rc) {
        release_staging(r->st);
        delete r;
        return rc;
    }
    r->live = true;
    h->resident.push_back(r);
    *dev = *host;
    dev->ref_read_off = r->view.ref_read_off;
    dev->starts = r->view.starts;
    dev->cigar_off = r->view.cigar_off;
    dev->cigar = r->view.cigar;
    dev->seq_woff = r->view.seq_woff;
    dev->planes = (const uint64_t *)r->view.planes;
    dev->okmask = r->view.okmask;
    dev->exc_read = r->view.exc_read;
    dev->exc_pos = r->view.exc_pos;
    dev->on_device = 1;
    dev->reserved = (uint32_t)h->resident.size();
    return BC_OK;
}

void bc_batch_free(bc_handle *h, bc_batch *dev)
{
    if (!h || !dev || !dev->on_device) return;
    const uint32_t id = dev->reserved;
    if (id == 0 || id > h->resident.size() || !h->resident[id - 1]) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    release_staging(h->resident[id - 1]->st);
    delete h->resident[id - 1];
    h->resident[id - 1] = nullptr;
    dev->reserved = 0;
}

// ------------------------------------------------------------------ results
int bc_counts(bc_handle *h, uint32_t ref, int64_t *out)
{
    if (!h || !out) return BC_ERR_ARG;
    if (ref >= h->n_refs) return fail(h, BC_ERR_ARG, "reference slot out of range");
    CU(h, cudaSetDevice(h->device));
    const uint64_t n = (uint64_t)h->ref_len[ref] * kPlanes;
    if (n == 0) return BC_OK;
    int rc = ensure(h, h->scratch_i64, n * sizeof(long long));
    if (rc) return rc;
    k_export_counts<<<(unsigned)((n + 255) / 256), 256, 0, h->compute>>>(h->d_counts, h->d_counts64, h->stride,
                                                                        h->col_base[ref], h->ref_len[ref],
                                                                        (long long *)h->scratch_i64.p);
    h->launches++;
    CU(h, cudaMemcpyAsync(out, h->scratch_i64.p, n * sizeof(long long), cudaMemcpyDeviceToHost, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

static int run_rows(bc_handle *h, uint32_t ref, int K, double norm, double norm2, bool want_cov, bool want_pc,
                    bool want_ent, bool want_sec, bool want_flags, uint32_t lo = 0, uint32_t n = 0xFFFFFFFFu)
{
    const uint32_t L = n == 0xFFFFFFFFu ? h->ref_len[ref] : n;      // a window [lo, lo + n) of the slot, or all of it
    int rc;
    if (want_cov && (rc = ensure(h, h->scratch_cov, (size_t)L * 8))) return rc;
    if (want_pc && (rc = ensure(h, h->scratch_pc, (size_t)L * 8 * K))) return rc;
    if (want_ent && (rc = ensure(h, h->scratch_ent, (size_t)L * 8))) return rc;
    if (want_sec && (rc = ensure(h, h->scratch_sec, (size_t)L * 8))) return rc;
    if (want_flags && (rc = ensure(h, h->scratch_flags, (size_t)L))) return rc;
    k2_stats_rows<<<(L + 255) / 256, 256, 0, h->compute>>>(
        h->d_counts, h->d_counts64, h->stride, (uint64_t)h->col_base[ref] + lo, L, K, norm, norm2,
        want_cov ? (long long *)h->scratch_cov.p : nullptr, want_pc ? (double *)h->scratch_pc.p : nullptr,
        want_ent ? (double *)h->scratch_ent.p : nullptr, want_sec ? (double *)h->scratch_sec.p : nullptr,
        want_flags ? (uint8_t *)h->scratch_flags.p : nullptr);
    h->launches++;
    CU(h, cudaGetLastError());
    return BC_OK;
}

int bc_stats(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, int64_t *coverage, double *pc,
             double *entropy, double *secondary, uint8_t *flags)
{
    if (!h) return BC_ERR_ARG;
    if (ref >= h->n_refs) return fail(h, BC_ERR_ARG, "reference slot out of range");
    CU(h, cudaSetDevice(h->device));
    const uint32_t L = h->ref_len[ref];
    if (L == 0) return BC_OK;
    const int K = show_n ? 6 : 5;
    int rc = run_rows(h, ref, K, norm, norm2, coverage, pc, entropy, secondary, flags);
    if (rc) return rc;
    if (coverage) CU(h, cudaMemcpyAsync(coverage, h->scratch_cov.p, (size_t)L * 8, cudaMemcpyDeviceToHost, h->compute));
    if (pc) CU(h, cudaMemcpyAsync(pc, h->scratch_pc.p, (size_t)L * 8 * K, cudaMemcpyDeviceToHost, h->compute));
    if (entropy) CU(h, cudaMemcpyAsync(entropy, h->scratch_ent.p, (size_t)L * 8, cudaMemcpyDeviceToHost, h->compute));
    if (secondary) CU(h, cudaMemcpyAsync(secondary, h->scratch_sec.p, (size_t)L * 8, cudaMemcpyDeviceToHost, h->compute));
    if (flags) CU(h, cudaMemcpyAsync(flags, h->scratch_flags.p, (size_t)L, cudaMemcpyDeviceToHost, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

int bc_rows_window(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t lo, uint32_t n,
                   int64_t *counts, int64_t *coverage, double *pc, double *entropy, double *secondary, uint8_t *flags)
{
    if (!h) return BC_ERR_ARG;
    if (ref >= h->n_refs) return fail(h, BC_ERR_ARG, "reference slot out of range");
    if ((uint64_t)lo + n > h->ref_len[ref]) return fail(h, BC_ERR_ARG, "bc_rows_window: window past the end of the slot");
    if (n == 0) return BC_OK;
    CU(h, cudaSetDevice(h->device));
    const int K = show_n ? 6 : 5;
    int rc;
    if (counts) {
        const uint64_t cells = (uint64_t)n * kPlanes;
        if ((rc = ensure(h, h->scratch_i64, cells * sizeof(long long)))) return rc;
        k_export_counts<<<(unsigned)((cells + 255) / 256), 256, 0, h->compute>>>(
            h->d_counts, h->d_counts64, h->stride, (uint64_t)h->col_base[ref] + lo, n, (long long *)h->scratch_i64.p);
        h->launches++;
        CU(h, cudaMemcpyAsync(counts, h->scratch_i64.p, cells * sizeof(long long), cudaMemcpyDeviceToHost, h->compute));
    }
    if (coverage || pc || entropy || secondary || flags) {
        if ((rc = run_rows(h, ref, K, norm, norm2, coverage, pc, entropy, secondary, flags, lo, n))) return rc;
        if (coverage) CU(h, cudaMemcpyAsync(coverage, h->scratch_cov.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->compute));
        if (pc) CU(h, cudaMemcpyAsync(pc, h->scratch_pc.p, (size_t)n * 8 * K, cudaMemcpyDeviceToHost, h->compute));
        if (entropy) CU(h, cudaMemcpyAsync(entropy, h->scratch_ent.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->compute));
        if (secondary) CU(h, cudaMemcpyAsync(secondary, h->scratch_sec.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->compute));
        if (flags) CU(h, cudaMemcpyAsync(flags, h->scratch_flags.p, (size_t)n, cudaMemcpyDeviceToHost, h->compute));
    }
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

static int summary_impl(bc_handle *h, int show_n, double norm, double norm2, int64_t *nonzero, int64_t *cov_sum,
                        double *entropy_sum, bool sync, long long min_cov = -1, bool allreduce = false);

int bc_summary(bc_handle *h, int show_n, double norm, double norm2, int64_t *nonzero, int64_t *cov_sum,
               double *entropy_sum)
{
    return summary_impl(h, show_n, norm, norm2, nonzero, cov_sum, entropy_sum, true);
}

int bc_summary_async(bc_handle *h, int show_n, double norm, double norm2, int64_t *nonzero, int64_t *cov_sum,
                     double *entropy_sum)
{
    return summary_impl(h, show_n, norm, norm2, nonzero, cov_sum, entropy_sum, false);
}

int bc_summary_min_coverage(bc_handle *h, int show_n, double norm, int64_t min_coverage, int64_t *selected,
                            int64_t *cov_sum, double *entropy_sum_selected)
{
    return summary_impl(h, show_n, norm, 0.0, selected, cov_sum, entropy_sum_selected, true,
                        min_coverage < 0 ? 0 : (long long)min_coverage);
}

int bc_summary_allreduce_async(bc_handle *h, int show_n, double norm, double norm2, int64_t *nonzero, int64_t *cov_sum,
                               double *entropy_sum)
{
    if (!h) return BC_ERR_ARG;
    if (!h->comm) return fail(h, BC_ERR_STATE, "bc_summary_allreduce_async: bc_comm_init has not been called");
    return summary_impl(h, show_n, norm, norm2, nonzero, cov_sum, entropy_sum, false, -1, true);
}

static int summary_impl(bc_handle *h, int show_n, double norm, double norm2, int64_t *nonzero, int64_t *cov_sum,
                        double *entropy_sum, bool sync, long long min_cov, bool allreduce)
{
    if (!h || !nonzero || !cov_sum || !entropy_sum) return BC_ERR_ARG;
    if (h->n_refs == 0) return fail(h, BC_ERR_STATE, "bc_begin has not been called");
    CU(h, cudaSetDevice(h->device));
    const uint32_t R = h->n_refs;
    // partial offsets per slot (a fixed function of the slot lengths), uploaded once per bc_begin
    if (h->part_off_refs != R || !h->d_part_off) {
        std::vector<uint32_t> off(R);
        size_t need = 0;
        uint32_t maxb = 1;
        for (uint32_t r = 0; r < R; r++) {
            off[r] = (uint32_t)need;
            const uint32_t nb = summary_blocks(h->slot_cap[r]);    // (room for the slot at its longest: bc_set_length)
            need += nb;
            maxb = std::max(maxb, nb);
        }
        CU(h, cudaStreamSynchronize(h->compute));         // earlier summaries still read the old offsets
        CU(h, cudaStreamSynchronize(h->stats));
        if (need > h->partials_cap || R != h->part_off_cap || !h->d_part_off) {
            if (h->d_partials) CU(h, cudaFree(h->d_partials));
            if (h->d_part_off) CU(h, cudaFree(h->d_part_off));
            h->d_partials = nullptr;
            h->d_part_off = nullptr;
            CU(h, cudaMalloc(&h->d_partials, need * sizeof(SummaryPartial)));
            CU(h, cudaMalloc(&h->d_part_off, (size_t)R * 2 * sizeof(uint32_t)));   // offsets, then arrival counters
            CU(h, cudaMemset(h->d_part_off, 0, (size_t)R * 2 * sizeof(uint32_t)));
            h->partials_cap = need;
            h->part_off_cap = R;
        }
        CU(h, cudaMemcpy(h->d_part_off, off.data(), (size_t)R * sizeof(uint32_t), cudaMemcpyHostToDevice));
        h->part_off_host = off;
        h->total_partials = need;
        h->part_off_refs = R;
        h->summary_max_blocks = maxb;
    }
    const int K = show_n ? 6 : 5;
    (void)norm2;                                          // --summarise does not need the secondary entropy
    const dim3 grid(h->summary_max_blocks, R);
    size_t off = 0;
    if (!allreduce) {
        // per-CTA partials straight into the result arena; the host adds them up when it fetches the arena
        int rcr = reserve_results(h, h->total_partials * sizeof(SummaryPartial), &off);
        if (rcr) return rcr;
        SummaryPartial *d_part = reinterpret_cast<SummaryPartial *>(h->d_results + off);
        // asynchronous: on the summary stream, behind everything the compute stream holds so far (see bc_handle::stats)
        cudaStream_t sstream = h->compute;
        if (!sync && h->stats_overlap) {
            sstream = h->stats;
            CU(h, cudaEventRecord(h->stats_ready, h->compute));
            CU(h, cudaStreamWaitEvent(h->stats, h->stats_ready, 0));
        }
        if (h->d_counts64)
            k2_summary<true, false><<<grid, 256, 0, sstream>>>(h->d_counts, h->d_counts64, h->stride, h->d_col_base, h->d_ref_len,
                                                               K, norm, min_cov, h->d_log2_tab, h->d_part_off, d_part, nullptr,
                                                               nullptr, nullptr, nullptr);
        else
            k2_summary<false, false><<<grid, 256, 0, sstream>>>(h->d_counts, nullptr, h->stride, h->d_col_base, h->d_ref_len, K,
                                                                norm, min_cov, h->d_log2_tab, h->d_part_off, d_part, nullptr,
                                                                nullptr, nullptr, nullptr);
        if (sstream == h->stats) {
            CU(h, cudaEventRecord(h->stats_done, h->stats));
            h->stats_busy = true;
            h->stats_set = h->cur_set;
        }
        h->launches += 1;
        bc_handle::PendingReduce q;
        q.off = off;
        q.first = h->part_off_host;
        q.count.resize(R);
        for (uint32_t r = 0; r < R; r++) q.count[r] = summary_blocks(h->ref_len[r]);
        q.nz = nonzero;
        q.cs = cov_sum;
        q.es = entropy_sum;
        h->pending_reduce.push_back(std::move(q));
    } else {
        int rcr = reserve_results(h, (size_t)R * 24, &off);
        if (rcr) return rcr;
        long long *d_nz = (long long *)(h->d_results + off);   // device arena, fetched at the next synchronisation
        long long *d_cs = d_nz + R;
        double *d_es = (double *)(d_cs + R);
        if (h->d_counts64)
            k2_summary<true, true><<<grid, 256, 0, h->compute>>>(h->d_counts, h->d_counts64, h->stride, h->d_col_base, h->d_ref_len, K,
                                                                 norm, min_cov, h->d_log2_tab, h->d_part_off, h->d_partials,
                                                                 h->d_part_off + R, d_nz, d_cs, d_es);
        else
            k2_summary<false, true><<<grid, 256, 0, h->compute>>>(h->d_counts, nullptr, h->stride, h->d_col_base, h->d_ref_len, K, norm,
                                                                  min_cov, h->d_log2_tab, h->d_part_off, h->d_partials,
                                                                  h->d_part_off + R, d_nz, d_cs, d_es);
        h->launches += 1;
        if (h->comm_world > 1) {
            // every rank holds the sums over the columns it owns: the reference's numbers are the sums over ranks
            // (main.py:479-485).  nonzero and cov_sum are adjacent int64 arrays; entropy sums are float64.
            // The two reductions are one NCCL group: aggregated into a single launch (one collective's latency per step).
            bcnccl::Api &nc = bcnccl::api();
            ncclResult_t r0 = nc.GroupStart();
            ncclResult_t r1 = r0 == ncclSuccess ? nc.AllReduce(d_nz, d_nz, (size_t)R * 2, ncclInt64, ncclSum, h->comm, h->compute) : r0;
            ncclResult_t r2 = r1 == ncclSuccess ? nc.AllReduce(d_es, d_es, (size_t)R, ncclFloat64, ncclSum, h->comm, h->compute) : r1;
            ncclResult_t r3 = nc.GroupEnd();
            if (r2 == ncclSuccess) r2 = r3;
            if (r2 != ncclSuccess) return fail(h, BC_ERR_CUDA, nc.GetErrorString(r2));
        }
        h->pending.push_back({off, (size_t)R * 8, nonzero});
        h->pending.push_back({off + (size_t)R * 8, (size_t)R * 8, cov_sum});
        h->pending.push_back({off + (size_t)R * 16, (size_t)R * 8, entropy_sum});
    }
    CU(h, cudaGetLastError());
    if (sync) {
        int rcf = fetch_summaries(h);
        if (rcf) return rcf;
        CU(h, cudaStreamSynchronize(h->compute));
        deliver_summaries(h);
    }
    return BC_OK;
}

static int amplicons_impl(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t n_tiles,
                          const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty, bool sync);

int bc_amplicons(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t n_tiles,
                 const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty)
{
    return amplicons_impl(h, ref, show_n, norm, norm2, n_tiles, lo, hi, out, empty, true);
}

int bc_amplicons_async(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t n_tiles,
                       const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty)
{
    return amplicons_impl(h, ref, show_n, norm, norm2, n_tiles, lo, hi, out, empty, false);
}

static int amplicons_impl(bc_handle *h, uint32_t ref, int show_n, double norm, double norm2, uint32_t n_tiles,
                          const int32_t *lo, const int32_t *hi, double *out, uint8_t *empty, bool sync)
{
    if (!h) return BC_ERR_ARG;
    if (ref >= h->n_refs) return fail(h, BC_ERR_ARG, "reference slot out of range");
    if (n_tiles == 0) return BC_OK;
    if (!lo || !hi || !out || !empty) return fail(h, BC_ERR_ARG, "bc_amplicons: null argument");
    CU(h, cudaSetDevice(h->device));
    const uint32_t L = h->ref_len[ref];
    const int K = show_n ? 6 : 5;
    int rc = BC_OK;
    if (L) rc = run_rows(h, ref, K, norm, norm2, true, false, true, true, false);
    if (rc) return rc;
    // results in the arena (fetched at the next synchronisation): out[6T] empty[T]; device scratch: lo[T] hi[T]
    const size_t out_bytes = ((size_t)n_tiles * 49 + 7) / 8 * 8;
    size_t off = 0;
    if ((rc = reserve_results(h, out_bytes, &off))) return rc;
    double *d_out = (double *)(h->d_results + off);
    uint8_t *d_empty = (uint8_t *)(h->d_results + off + (size_t)n_tiles * 48);
    const bool same_tiles = h->tiles_dev.p && h->tiles_host.size() == (size_t)n_tiles * 2 &&
                            std::memcmp(h->tiles_host.data(), lo, (size_t)n_tiles * 4) == 0 &&
                            std::memcmp(h->tiles_host.data() + n_tiles, hi, (size_t)n_tiles * 4) == 0;
    if (!same_tiles) {
        // (earlier launches may still read the old windows: the copy is ordered behind them on the same stream, from a
        //  host copy that lives as long as the handle -- the caller's arrays are done with when this call returns)
        if ((rc = ensure(h, h->tiles_dev, (size_t)n_tiles * 8 + 64))) return rc;
        CU(h, cudaStreamSynchronize(h->compute));
        h->tiles_host.assign(lo, lo + n_tiles);
        h->tiles_host.insert(h->tiles_host.end(), hi, hi + n_tiles);
        CU(h, cudaMemcpyAsync(h->tiles_dev.p, h->tiles_host.data(), (size_t)n_tiles * 8, cudaMemcpyHostToDevice, h->compute));
    }
    int32_t *d_lo = (int32_t *)h->tiles_dev.p;
    int32_t *d_hi = d_lo + n_tiles;
    uint32_t cap = 4096;                        // doubles staged per window (32 KB)
    k3_amplicons<<<dim3(n_tiles, 3), kK3Threads, (size_t)cap * sizeof(double), h->compute>>>(
        (const long long *)h->scratch_cov.p, (const double *)h->scratch_ent.p, (const double *)h->scratch_sec.p, L,
        d_lo, d_hi, n_tiles, cap, d_out, d_empty);
    h->launches++;
    CU(h, cudaGetLastError());
    h->pending.push_back({off, (size_t)n_tiles * 48, out});
    h->pending.push_back({off + (size_t)n_tiles * 48, (size_t)n_tiles, empty});
    if (sync) {
        if ((rc = fetch_summaries(h))) return rc;
        CU(h, cudaStreamSynchronize(h->compute));
        deliver_summaries(h);
    }
    return BC_OK;
}

// ------------------------------------------------------------------ halo exchange (region sharding)
__global__ void k_halo_export(const uint32_t *__restrict__ c32, uint64_t stride, uint64_t col, uint32_t n,
                              uint32_t *__restrict__ buf)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * kPlanes) return;
    buf[t] = c32[(uint64_t)(t / n) * stride + col + (t % n)];
}
__global__ void k_halo_add(uint32_t *__restrict__ c32, uint64_t stride, uint64_t col, uint32_t n,
                           const uint32_t *__restrict__ buf)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * kPlanes) return;
    c32[(uint64_t)(t / n) * stride + col + (t % n)] += buf[t];
}

int bc_halo_export(bc_handle *h, uint32_t ref, uint32_t col_lo, uint32_t n_cols, uint32_t *dev_buf)
{
    if (!h || !dev_buf) return BC_ERR_ARG;
    if (ref >= h->n_refs || (uint64_t)col_lo + n_cols > h->ref_len[ref]) return fail(h, BC_ERR_ARG, "halo out of range");
    if (h->d_counts64) return fail(h, BC_ERR_STATE, "bc_halo_export: accumulators were folded to int64 (more than 2^32 reads since bc_begin)");
    if (n_cols == 0) return BC_OK;
    CU(h, cudaSetDevice(h->device));
    h->side_needs_compute = true;
    k_halo_export<<<(n_cols * kPlanes + 255) / 256, 256, 0, h->compute>>>(h->d_counts, h->stride,
                                                                         (uint64_t)h->col_base[ref] + col_lo, n_cols, dev_buf);
    h->launches++;
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

int bc_halo_add(bc_handle *h, uint32_t ref, uint32_t col_lo, uint32_t n_cols, const uint32_t *dev_buf)
{
    if (!h || !dev_buf) return BC_ERR_ARG;
    if (ref >= h->n_refs || (uint64_t)col_lo + n_cols > h->ref_len[ref]) return fail(h, BC_ERR_ARG, "halo out of range");
    if (h->d_counts64) return fail(h, BC_ERR_STATE, "bc_halo_add: accumulators were folded to int64 (more than 2^32 reads since bc_begin)");
    if (n_cols == 0) return BC_OK;
    CU(h, cudaSetDevice(h->device));
    JOIN_STATS(h);
    h->side_needs_compute = true;
    k_halo_add<<<(n_cols * kPlanes + 255) / 256, 256, 0, h->compute>>>(h->d_counts, h->stride,
                                                                      (uint64_t)h->col_base[ref] + col_lo, n_cols, dev_buf);
    h->launches++;
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

int bc_truncate(bc_handle *h, uint32_t ref, uint32_t new_len)
{
    if (!h) return BC_ERR_ARG;
    if (ref >= h->n_refs || new_len > h->slot_cap[ref]) return fail(h, BC_ERR_ARG, "bc_truncate: out of range");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->side));               // an overflow check may still read the old length
    JOIN_STATS(h);                                       // ... and so may a summary on its own stream
    h->ref_len[ref] = new_len;
    h->side_needs_compute = true;
    h->part_off_refs = 0;
    CU(h, cudaMemcpyAsync(h->d_ref_len + ref, &h->ref_len[ref], sizeof(uint32_t), cudaMemcpyHostToDevice, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

int bc_set_length(bc_handle *h, uint32_t ref, uint32_t new_len)
{
    if (!h) return BC_ERR_ARG;
    if (ref >= h->n_refs || new_len > h->slot_cap[ref]) return fail(h, BC_ERR_ARG, "bc_set_length: out of range");
    CU(h, cudaSetDevice(h->device));
    // the overflow check of the last batch (side stream) may still read the old length: order the write behind it
    CU(h, cudaStreamWaitEvent(h->compute, h->checked, 0));
    JOIN_STATS(h);                                       // ... and so may a summary on its own stream
    h->ref_len[ref] = new_len;
    h->side_needs_compute = true;
    k_set_u32<<<1, 1, 0, h->compute>>>(h->d_ref_len + ref, new_len);
    h->launches++;
    CU(h, cudaGetLastError());
    return BC_OK;
}

// ---- the communicator ---------------------------------------------------------------------------------------
static int nccl_fail(bc_handle *h, ncclResult_t r, const char *what)
{
    bcnccl::Api &nc = bcnccl::api();
    h->err = std::string(what) + ": " + (nc.GetErrorString ? nc.GetErrorString(r) : "NCCL error");
    return BC_ERR_CUDA;
}

int bc_comm_unique_id(void *id128)
{
    if (!id128) return BC_ERR_ARG;
    bcnccl::Api &nc = bcnccl::api();
    if (!nc.ok()) {
        g_create_error = nc.err;
        return BC_ERR_CUDA;
    }
    ncclUniqueId id;
    if (nc.GetUniqueId(&id) != ncclSuccess) {
        g_create_error = "ncclGetUniqueId failed";
        return BC_ERR_CUDA;
    }
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, sizeof(id));
    return BC_OK;
}

int bc_comm_init(bc_handle *h, int world, int rank, const void *id128)
{
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return BC_ERR_ARG;
    bcnccl::Api &nc = bcnccl::api();
    if (!nc.ok()) return fail(h, BC_ERR_CUDA, nc.err.c_str());
    if (h->comm) return fail(h, BC_ERR_STATE, "bc_comm_init: the handle already has a communicator");
    CU(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = nc.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) {
        h->comm = nullptr;
        return nccl_fail(h, r, "ncclCommInitRank");
    }
    h->comm_world = world;
    h->comm_rank = rank;
    return BC_OK;
}

int bc_comm_destroy(bc_handle *h)
{
    if (!h) return BC_ERR_ARG;
    if (h->comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->compute);
        bcnccl::api().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->comm_world = 1;
    h->comm_rank = 0;
    return BC_OK;
}

int bc_comm_allgather_u32(bc_handle *h, uint32_t mine, uint32_t *out)
{
    if (!h || !out) return BC_ERR_ARG;
    if (!h->comm) return fail(h, BC_ERR_STATE, "bc_comm_allgather_u32: bc_comm_init has not been called");
    CU(h, cudaSetDevice(h->device));
    const int W = h->comm_world;
    int rc = ensure(h, h->comm_scratch, (size_t)(W + 1) * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t *d = (uint32_t *)h->comm_scratch.p;
    k_set_u32<<<1, 1, 0, h->compute>>>(d + W, mine);
    ncclResult_t r = bcnccl::api().AllGather(d + W, d, 1, ncclUint32, h->comm, h->compute);
    if (r != ncclSuccess) return nccl_fail(h, r, "ncclAllGather");
    CU(h, cudaMemcpyAsync(out, d, (size_t)W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->compute));
    CU(h, cudaStreamSynchronize(h->compute));
    return BC_OK;
}

int bc_halo_merge(bc_handle *h, uint32_t ref, const uint32_t *bounds, const uint32_t *halos)
{
    if (!h || !bounds || !halos) return BC_ERR_ARG;
    if (!h->comm) return fail(h, BC_ERR_STATE, "bc_halo_merge: bc_comm_init has not been called");
    if (ref >= h->n_refs) return fail(h, BC_ERR_ARG, "bc_halo_merge: slot out of range");
    const int W = h->comm_world, me = h->comm_rank;
    for (int r = 0; r < W; r++)
        if (bounds[r] > bounds[r + 1]) return fail(h, BC_ERR_ARG, "bc_halo_merge: boundaries must not decrease");
    const uint64_t lo = bounds[me], hi = bounds[me + 1], own = hi - lo;
    if (own + halos[me] != h->ref_len[ref] || own + halos[me] > h->slot_cap[ref])
        return fail(h, BC_ERR_ARG, "bc_halo_merge: the slot must hold the owned columns plus this rank's halo");
    if (h->d_counts64) return fail(h, BC_ERR_STATE, "bc_halo_merge: accumulators were folded to int64");
    CU(h, cudaSetDevice(h->device));
    JOIN_STATS(h);
    struct Seg { int peer; uint64_t col; uint64_t n; };
    std::vector<Seg> sends, recvs;
    // what this rank sends: its halo covers global columns [hi, hi + halos[me])
    for (int s = me + 1; s < W; s++) {
        const uint64_t a = std::max<uint64_t>(bounds[s], hi), b = std::min<uint64_t>(bounds[s + 1], hi + halos[me]);
        if (a < b) sends.push_back({s, a - lo, b - a});
    }
    // what it receives: halos of ranks to the left that reach into [lo, hi)
    uint64_t recv_cols = 0;
    for (int q = 0; q < me; q++) {
        const uint64_t hq = bounds[q + 1];
        const uint64_t a = std::max<uint64_t>(lo, hq), b = std::min<uint64_t>(hi, hq + halos[q]);
        if (a < b) {
            recvs.push_back({q, a - lo, b - a});
            recv_cols += b - a;
        }
    }
    int rc = ensure(h, h->halo_recv, (size_t)std::max<uint64_t>(recv_cols, 1) * kPlanes * sizeof(uint32_t));
    if (rc) return rc;
    bcnccl::Api &nc = bcnccl::api();
    const uint64_t base = h->col_base[ref];
    if (!sends.empty() || !recvs.empty()) {
        // plane segments leave straight from the accumulators (K1 and its corrections are ahead on this stream)
        ncclResult_t r = nc.GroupStart();
        if (r != ncclSuccess) return nccl_fail(h, r, "ncclGroupStart");
        for (const Seg &sg : sends)
            for (int p = 0; p < kPlanes && r == ncclSuccess; p++)
                r = nc.Send(h->d_counts + (uint64_t)p * h->stride + base + sg.col, sg.n, ncclUint32, sg.peer, h->comm, h->compute);
        uint64_t at = 0;
        for (const Seg &sg : recvs) {
            for (int p = 0; p < kPlanes && r == ncclSuccess; p++)
                r = nc.Recv((uint32_t *)h->halo_recv.p + at + (uint64_t)p * sg.n, sg.n, ncclUint32, sg.peer, h->comm, h->compute);
            at += sg.n * kPlanes;
        }
        ncclResult_t re = nc.GroupEnd();
        if (r != ncclSuccess || re != ncclSuccess) return nccl_fail(h, r != ncclSuccess ? r : re, "halo send/recv");
        at = 0;
        for (const Seg &sg : recvs) {
            k_halo_add<<<(unsigned)((sg.n * kPlanes + 255) / 256), 256, 0, h->compute>>>(h->d_counts, h->stride, base + sg.col,
                                                                                      (uint32_t)sg.n, (const uint32_t *)h->halo_recv.p + at);
            h->launches++;
            at += sg.n * kPlanes;
        }
    }
    return bc_set_length(h, ref, (uint32_t)own);
}

// ------------------------------------------------------------------ instrumentation
int bc_timer_start(bc_handle *h)
{
    if (!h) return BC_ERR_ARG;
    CU(h, cudaEventRecord(h->t0, h->compute));
    return BC_OK;
}

int bc_timer_stop(bc_handle *h, float *ms)
{
    if (!h || !ms) return BC_ERR_ARG;
    CU(h, cudaStreamWaitEvent(h->compute, h->checked, 0));   // work forked to the side stream belongs to the timed region
    JOIN_STATS(h);                                           // ... and so do the summaries on theirs
    CU(h, cudaEventRecord(h->t1, h->compute));
    CU(h, cudaEventSynchronize(h->t1));
    CU(h, cudaEventElapsedTime(ms, h->t0, h->t1));
    return BC_OK;
}

int bc_h2d_probe(bc_handle *h, uint64_t bytes, int reps, double *gb_per_s)
{
    if (!h || !gb_per_s || bytes == 0 || reps <= 0) return BC_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    void *host = nullptr, *dev = nullptr;
    CU(h, cudaHostAlloc(&host, bytes, cudaHostAllocDefault));
    if (cudaMalloc(&dev, bytes) != cudaSuccess) {
        cudaFreeHost(host);
        return fail(h, BC_ERR_CUDA, "bc_h2d_probe: cudaMalloc failed");
    }
    std::memset(host, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, h->copy);      // warm-up
    cudaEventRecord(a, h->copy);
    for (int i = 0; i < reps; i++) cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, h->copy);
    cudaEventRecord(b, h->copy);
    cudaError_t e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(dev);
    cudaFreeHost(host);
    if (e != cudaSuccess || ms <= 0.f) return fail(h, BC_ERR_CUDA, "bc_h2d_probe: copy failed");
    *gb_per_s = (double)bytes * reps / (ms * 1e-3) / 1e9;
    return BC_OK;
}

int bc_last_count_kernel_ms(bc_handle *h, float *ms)
{
    return bc_count_kernel_ms_history(h, ms, 1) == 1 ? BC_OK : BC_ERR_STATE;
}

int bc_count_kernel_ms_history(bc_handle *h, float *ms, int n)
{
    if (!h || !ms || n <= 0) return -1;
    const int have = (int)std::min<uint64_t>(h->k_count, (uint64_t)bc_handle::kHist);
    n = std::min(n, have);
    for (int i = 0; i < n; i++) {                 // ms[0] = most recent launch
        const int ki = (int)((h->k_count - 1 - (uint64_t)i) % bc_handle::kHist);
        if (cudaEventSynchronize(h->k1[ki]) != cudaSuccess) return -1;
        if (cudaEventElapsedTime(&ms[i], h->k0[ki], h->k1[ki]) != cudaSuccess) return -1;
    }
    return n;
}

uint64_t bc_kernel_launches(bc_handle *h) { return h ? h->launches : 0; }

int bc_set_count_variant(bc_handle *h, int variant)
{
    if (!h || variant < 0 || variant > 2) return BC_ERR_ARG;
    h->variant = variant == 0 ? h->default_variant : variant;
    for (Resident *r : h->resident)
        if (r) r->walker = 0;
    return BC_OK;
}

// ------------------------------------------------------------------ host packer
// The packers' CIGAR normal form (csrc/cigar_canon.h) for n reads: out_off[n + 1] and, when `out` is not NULL, the words.
// Returns the number of words of the normal form.
uint64_t bc_canonical_cigars(uint32_t n_reads, const uint32_t *cigar, const uint64_t *cigar_off, uint32_t *out, uint32_t *out_off)
{
    if (!cigar_off || !out_off || (n_reads && cigar_off[n_reads] && !cigar)) return ~0ull;
    uint64_t m = 0;
    out_off[0] = 0;
    for (uint32_t i = 0; i < n_reads; i++) {
        const uint32_t *src = cigar + cigar_off[i];
        m += bccanon::canon_cigar((uint32_t)(cigar_off[i + 1] - cigar_off[i]), [&](uint32_t k) { return src[k]; }, out ? out + m : nullptr);
        if (m > 0xFFFFFFFFull) return ~0ull;
        out_off[i + 1] = (uint32_t)m;
    }
    return m;
}

uint64_t bc_pack_words(uint32_t n_reads, const uint64_t *seq_off)
{
    uint64_t w = 0;
    for (uint32_t i = 0; i < n_reads; i++) w += (seq_off[i + 1] - seq_off[i] + 31) / 32;
    return w;
}

int bc_pack_reads(uint32_t n_reads, const uint8_t *seq, const uint8_t *qual, const uint64_t *seq_off,
                  const uint32_t *cigar, const uint32_t *cigar_off, uint32_t min_base_quality,
                  uint32_t *seq_woff_out, uint64_t *planes_out, uint32_t *okmask_out, uint32_t *exc_read_out,
                  uint32_t *exc_pos_out, uint32_t exc_cap, uint32_t *n_exc)
{
    if (!seq_off || !seq_woff_out || !n_exc) return BC_ERR_ARG;
    if (min_base_quality > 0 && (!okmask_out || !qual)) return BC_ERR_ARG;
    // letter class: 0..3 = A,C,G,T; 4 = N; 5 = anything else (never counted, count.cpp:58-65)
    uint8_t cls[256];
    std::memset(cls, 5, sizeof(cls));
    cls['A'] = 0; cls['C'] = 1; cls['G'] = 2; cls['T'] = 3; cls['N'] = 4;
    // word offsets first (every read starts on a fresh 32-base word), then the reads in parallel
    uint64_t w = 0;
    for (uint32_t i = 0; i < n_reads; i++) {
        const uint64_t len = seq_off[i + 1] - seq_off[i];
        if (w > 0xFFFFFFFFull || len >= (1ull << 30)) return BC_ERR_ARG;
        seq_woff_out[i] = (uint32_t)w;
        w += (len + 31) / 32;
    }
    if (w > 0xFFFFFFFFull) return BC_ERR_ARG;
    seq_woff_out[n_reads] = (uint32_t)w;

    const uint64_t grain = 2048;
    const uint64_t chunks = ((uint64_t)n_reads + grain - 1) / grain;
    std::vector<std::vector<uint32_t>> exc(chunks);              // per chunk: (read, pos << 2 | flags) pairs, in order
    std::atomic<int> overrun(0);
    const int threads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    bcbam::parallel_for(threads, n_reads, grain, [&](uint64_t a, uint64_t e) {
        std::vector<uint32_t> &ex = exc[a / grain];
        for (uint64_t i = a; i < e; i++) {
            const uint64_t s0 = seq_off[i], len = seq_off[i + 1] - s0;
            if (cigar && cigar_off) {      // count.cpp:56,58 index quals/read unchecked: reject what would be UB there
                uint64_t rp = 0;
                for (uint32_t c = cigar_off[i]; c < cigar_off[i + 1]; c++) {
                    const uint32_t op = cigar[c] & 0xFu, l = cigar[c] >> 4;
                    if (op == 0 || op == 7 || op == 8) {
                        if (l && rp + l > len) overrun = 1;
                        rp += l;
                    } else if (op == 1) {
                        rp += l;
                    }
                }
            }
            uint64_t wi = seq_woff_out[i];
            for (uint64_t j0 = 0; j0 < len; j0 += 32, wi++) {
                const uint32_t m = (uint32_t)std::min<uint64_t>(32, len - j0);
                uint32_t lo = 0, hi = 0, ok = 0, odd = 0;            // odd: bases outside ACGT (0.1 % of real data)
                const uint8_t *sp = seq + s0 + j0;
                const uint8_t *qp = qual ? qual + s0 + j0 : nullptr;
#if defined(__SSE2__)
                if (m == 32) {                                      // 32 bases per step: byte compares + movemask
                    uint32_t mc = 0, mg = 0, mt = 0, ma = 0;
                    for (int h = 0; h < 2; h++) {
                        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(sp + 16 * h));
                        ma |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8('A'))) << (16 * h);
                        mc |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8('C'))) << (16 * h);
                        mg |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8('G'))) << (16 * h);
                        mt |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8('T'))) << (16 * h);
                    }
                    lo = mc | mt;                                   // codes A 0, C 1, G 2, T 3
                    hi = mg | mt;
                    odd = ~(ma | mc | mg | mt);
                } else
#endif
                {
                    for (uint32_t j = 0; j < m; j++) {              // branch-free: exceptions are collected below
                        const uint32_t c = cls[sp[j]];
                        lo |= (c & 1u) << j;
                        hi |= ((c >> 1) & 1u) << j;
                        odd |= (c >> 2) << j;
                    }
                    lo &= ~odd;
                    hi &= ~odd;
                }
                if (qp && min_base_quality > 0) {
#if defined(__SSE2__)
                    if (m == 32 && min_base_quality <= 255u) {
                        const __m128i thr = _mm_set1_epi8((char)min_base_quality);
                        for (int h = 0; h < 2; h++) {
                            const __m128i q = _mm_loadu_si128(reinterpret_cast<const __m128i *>(qp + 16 * h));
                            ok |= (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_max_epu8(q, thr), q)) << (16 * h);   // q >= thr
                        }
                    } else
#endif
                    for (uint32_t j = 0; j < m; j++) ok |= (uint32_t)(qp[j] >= min_base_quality) << j;
                } else {
                    ok = m == 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
                }
                for (uint32_t rest = odd; rest; rest &= rest - 1u) {
                    const uint32_t j = (uint32_t)__builtin_ctz(rest);
                    const bool is_n = cls[sp[j]] == 4;
                    uint32_t flags = 0;
                    if (min_base_quality == 0) flags = is_n ? 3u : 2u;             // undo the 'A', maybe count N
                    else if (is_n && ((ok >> j) & 1u)) flags = 1u;                  // masked out already; count N
                    if (flags) {
                        ex.push_back((uint32_t)i);
                        ex.push_back((uint32_t)((j0 + j) << 2) | flags);
                    }
                }
                ok &= ~odd;
                if (planes_out) planes_out[wi] = (uint64_t)lo | ((uint64_t)hi << 32);
                if (okmask_out) okmask_out[wi] = ok;
            }
        }
    });
    uint64_t ne = 0;
    for (const auto &ex : exc) {
        for (size_t k = 0; k + 1 < ex.size(); k += 2, ne++) {
            if (ne < exc_cap && exc_read_out && exc_pos_out) {
                exc_read_out[ne] = ex[k];
                exc_pos_out[ne] = ex[k + 1];
            }
        }
    }
    if (ne > 0xFFFFFFFFull) return BC_ERR_ARG;
    *n_exc = (uint32_t)ne;
    return overrun ? BC_ERR_READ_OVERRUN : BC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ native BAM decode (bam_decode.h)
static thread_local std::string g_bam_err;

extern "C" {

uint32_t bc_bgzf_crc32(const uint8_t *data, uint64_t n) { return data || n == 0 ? bcbam::crc32_fast(data, (size_t)n) : 0u; }

int bc_inflate_raw(const uint8_t *in, uint64_t in_len, uint8_t *out, uint64_t out_len)
{
    if ((!in && in_len) || (!out && out_len)) return 0;
    std::vector<uint8_t> padded((size_t)in_len + 16, 0);            // the decoder may read 8 bytes past the stream
    if (in_len) std::memcpy(padded.data(), in, (size_t)in_len);
    std::unique_ptr<bcbam::FastInflater> fi(new bcbam::FastInflater());
    return fi->inflate(padded.data(), (size_t)in_len, out, (size_t)out_len) ? 1 : 0;
}

int bc_bam_open(const char *path, int threads, bc_bam **out)
{
    if (!path || !out) return BC_ERR_ARG;
    *out = nullptr;
    g_bam_err.clear();
    const int rc = bc_bam_open_impl(path, threads, out, g_bam_err);
    return rc == 0 ? BC_OK : BC_ERR_ARG;
}

int bc_bam_stream_open(const char *path, int threads, bc_bam_stream **out)
{
    if (!path || !out) return BC_ERR_ARG;
    *out = nullptr;
    g_bam_err.clear();
    return bc_bam_stream_open_impl(path, threads, out, g_bam_err) == 0 ? BC_OK : BC_ERR_ARG;
}

int bc_bam_stream_next(bc_bam_stream *s, uint64_t max_inflated_bytes, bc_bam **out)
{
    if (!s || !out) return BC_ERR_ARG;
    g_bam_err.clear();
    return bc_bam_stream_next_impl(s, max_inflated_bytes, out, g_bam_err) == 0 ? BC_OK : BC_ERR_ARG;
}

void bc_bam_stream_close(bc_bam_stream *s) { delete s; }

int bc_bam_index_build(const char *bam_path, const char *bai_path, int threads)
{
    if (!bam_path || !bai_path) return BC_ERR_ARG;
    g_bam_err.clear();
    return bc_bam_index_build_impl(bam_path, bai_path, threads, g_bam_err) == 0 ? BC_OK : BC_ERR_ARG;
}

int bc_bam_open_region(const char *bam_path, const char *bai_path, int32_t ref_id, int64_t beg, int64_t end, int threads,
                       bc_bam **out)
{
    if (!bam_path || !bai_path || !out) return BC_ERR_ARG;
    *out = nullptr;
    g_bam_err.clear();
    return bc_bam_open_region_impl(bam_path, bai_path, ref_id, beg, end, threads, out, g_bam_err) == 0 ? BC_OK : BC_ERR_ARG;
}

const char *bc_bam_last_error(void) { return g_bam_err.c_str(); }

void bc_bam_close(bc_bam *b) { delete b; }

uint64_t bc_bam_num_records(const bc_bam *b) { return b ? b->rec_off.size() - 1 : 0; }

uint32_t bc_bam_num_refs(const bc_bam *b) { return b ? (uint32_t)b->ref_names.size() : 0; }

const char *bc_bam_ref_name(const bc_bam *b, uint32_t i) { return (b && i < b->ref_names.size()) ? b->ref_names[i].c_str() : nullptr; }

uint32_t bc_bam_ref_len(const bc_bam *b, uint32_t i) { return (b && i < b->ref_lens.size()) ? b->ref_lens[i] : 0; }

int bc_bam_core(const bc_bam *b, int32_t *ref_id, int32_t *pos, uint8_t *mapq, uint16_t *flag)
{
    if (!b) return BC_ERR_ARG;
    bc_bam_core_impl(b, ref_id, pos, mapq, flag);
    return BC_OK;
}

int bc_bam_select_sizes(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                        uint64_t *n_reads, uint64_t *n_cigar, uint64_t *n_bases)
{
    if (!b || !n_reads || !n_cigar || !n_bases || rec_a > rec_b || rec_b > b->rec_off.size() - 1) return BC_ERR_ARG;
    return bc_bam_select_sizes_impl(b, rec_a, rec_b, ref_id, min_mapq, n_reads, n_cigar, n_bases) ? BC_OK : BC_ERR_MISSING_QUAL;
}

int bc_bam_select_fill(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                       uint32_t *starts, uint32_t *cigar, uint64_t *cigar_off, uint8_t *seq, uint8_t *qual,
                       uint64_t *seq_off)
{
    if (!b || !cigar_off || !seq_off || rec_a > rec_b || rec_b > b->rec_off.size() - 1) return BC_ERR_ARG;
    bc_bam_select_fill_impl(b, rec_a, rec_b, ref_id, min_mapq, starts, cigar, cigar_off, seq, qual, seq_off);
    return BC_OK;
}

int bc_bam_pack_sizes(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq, uint64_t *out6)
{
    if (!b || !out6 || rec_a > rec_b || rec_b > b->rec_off.size() - 1) return BC_ERR_ARG;
    bc_pack_sizes z;
    bc_bam_pack_sizes_impl(b, rec_a, rec_b, ref_id, min_mapq, &z);
    out6[0] = z.n_reads;
    out6[1] = z.n_cigar;
    out6[2] = z.n_words;
    out6[3] = z.n_bases;
    out6[4] = z.aligned_bases;
    out6[5] = z.sorted;
    return z.missing_qual ? BC_ERR_MISSING_QUAL : BC_OK;
}

int bc_bam_pack_fill(const bc_bam *b, uint64_t rec_a, uint64_t rec_b, int32_t ref_id, uint32_t min_mapq,
                     uint32_t min_base_quality, uint32_t *starts, uint32_t *cigar, uint32_t *cigar_off, uint32_t *seq_woff,
                     uint64_t *planes, uint32_t *okmask, uint32_t *exc_read, uint32_t *exc_pos, uint64_t exc_cap,
                     uint64_t *n_exc)
{
    if (!b || !cigar_off || !seq_woff || !n_exc || rec_a > rec_b || rec_b > b->rec_off.size() - 1) return BC_ERR_ARG;
    if (min_base_quality > 0 && !okmask) return BC_ERR_ARG;
    const int rc = bc_bam_pack_fill_impl(b, rec_a, rec_b, ref_id, min_mapq, min_base_quality, starts, cigar, cigar_off,
                                         seq_woff, planes, okmask, exc_read, exc_pos, exc_cap, n_exc);
    return rc == 0 ? BC_OK : (rc == 5 ? BC_ERR_READ_OVERRUN : BC_ERR_ARG);
}

}  // extern "C"

// ------------------------------------------------------------------ exact TSV rows (tsv_format.h)
extern "C" {

int bc_format_tsv(const char *ref_name, uint64_t n_pos, uint64_t first_pos, int k, int long_format, int decimal_places,
                  const int64_t *counts, const int64_t *coverage, const double *pc, uint64_t pc_stride,
                  const double *entropy, const double *secondary, const uint8_t *flags, int threads, char **text,
                  uint64_t *len)
{
    if (!ref_name || !text || !len || (n_pos && (!counts || !coverage || !pc || !entropy || !secondary || !flags)))
        return BC_ERR_ARG;
    *text = nullptr;
    *len = 0;
    return bc_format_tsv_impl(ref_name, n_pos, first_pos, k, long_format, decimal_places, counts, coverage, pc, pc_stride,
                              entropy, secondary, flags, threads, text, len) == 0
               ? BC_OK
               : BC_ERR_ARG;
}

void bc_free_text(char *text) { std::free(text); }

}  // extern "C"
