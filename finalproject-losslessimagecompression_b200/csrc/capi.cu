// capi.cu -- the C ABI declared in include/flic_b200.h.
//
// Thin: argument checks, workspace carving, launches.  The host entry points add the
// host<->device copies, chunked over three CUDA streams so that the upload of one chunk, the
// kernels of the previous one and the download of the one before overlap, and synchronise once
// before returning.
#include "../../include/flic_b200.h"
#include "flic_core.cuh"
#include "flic_kernels.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

#define FLIC_CUDA(call)                                         \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
    } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Workspace layout for one encode call: [scratch u32 x n_symbols][counts i64 x n_streams][scan tmp]
struct EncodeWorkspace {
    uint32_t* scratch;
    int64_t* counts;
    int64_t* scan_tmp;
    int64_t bytes;
};

EncodeWorkspace carve(void* base, int64_t n_symbols, int64_t n_streams) {
    EncodeWorkspace w;
    char* p = (char*)base;
    int64_t off = 0;
    w.scratch = (uint32_t*)(p + off);
    off = align_up(off + (int64_t)sizeof(uint32_t) * (n_symbols > 0 ? n_symbols : 1), 256);
    w.counts = (int64_t*)(p + off);
    off = align_up(off + (int64_t)sizeof(int64_t) * (n_streams > 0 ? n_streams : 1), 256);
    w.scan_tmp = (int64_t*)(p + off);
    off = align_up(off + (int64_t)sizeof(int64_t) * flic::scan_tmp_elems(n_streams > 0 ? n_streams : 1), 256);
    w.bytes = off;
    return w;
}

}  // namespace

extern "C" {

int flic_abi_version(void) { return FLIC_ABI_VERSION; }
const char* flic_last_error(void) { return g_err; }
int64_t flic_kernel_launches(void) { return g_launches.load(); }
const char* flic_last_coder_kernel(int which) { return flic::last_coder_kernel(which); }
int64_t flic_decode_cluster_capacity(int cluster) { return flic::coop_cluster_capacity(cluster); }
int flic_set_decode_kernel(int which) {
    const bool known = which == -1 || which == 0 || which == 1 || which == 2 || which == 4 || which == 8;
    return flic::set_decode_kernel(known ? which : -1);
}

int flic_cdf_tables(const float* x, const float* mean, const float* scale, int64_t n_symbols,
                    uint32_t* start, uint32_t* freq, int32_t* status_word,
                    flic_cuda_stream_t stream) {
    if (n_symbols < 0) return fail(FLIC_E_ARG, "n_symbols < 0");
    if (n_symbols == 0) return 0;
    if (!x || !mean || !scale || !start || !freq || !status_word) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_cdf_tables(x, mean, scale, n_symbols, start, freq, status_word, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_debug_expf(const float* x, float* y, int64_t n, flic_cuda_stream_t stream) {
    if (n < 0 || (n > 0 && (!x || !y))) return fail(FLIC_E_ARG, "bad argument");
    FLIC_CUDA(flic::launch_debug_expf(x, y, n, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_debug_part1(const float* arg, int32_t* y, int64_t n, flic_cuda_stream_t stream) {
    if (n < 0 || (n > 0 && (!arg || !y))) return fail(FLIC_E_ARG, "bad argument");
    FLIC_CUDA(flic::launch_debug_part1(arg, y, n, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_debug_div_check(int64_t n, uint64_t seed, int mode, uint64_t* mismatches, flic_cuda_stream_t stream) {
    if (n < 0 || !mismatches || (mode != 0 && mode != 1)) return fail(FLIC_E_ARG, "bad argument");
    FLIC_CUDA(flic::launch_debug_div_check(n, seed, mode, (unsigned long long*)mismatches, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_debug_push_check(int64_t n, uint64_t seed, uint64_t* mismatches, flic_cuda_stream_t stream) {
    if (n < 0 || !mismatches) return fail(FLIC_E_ARG, "bad argument");
    FLIC_CUDA(flic::launch_debug_push_check(n, seed, (unsigned long long*)mismatches, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int64_t flic_encode_workspace_bytes(int64_t n_symbols, int64_t n_streams) {
    return carve(nullptr, n_symbols, n_streams).bytes;
}

int flic_rans_encode(const float* x, const float* mean, const float* scale,
                     const int64_t* stream_offsets, int64_t n_streams, int64_t n_symbols,
                     const uint64_t* init_states, void* workspace, int64_t workspace_bytes,
                     uint32_t* packed, int64_t packed_capacity, int64_t* word_offsets,
                     uint64_t* final_states, int32_t* status, flic_cuda_stream_t stream) {
    if (n_streams < 0 || n_symbols < 0) return fail(FLIC_E_ARG, "negative size");
    if (!word_offsets) return fail(FLIC_E_ARG, "null word_offsets");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_streams == 0) {
        FLIC_CUDA(cudaMemsetAsync(word_offsets, 0, sizeof(int64_t), st));
        return 0;
    }
    if (!stream_offsets || !workspace || !final_states || !status) return fail(FLIC_E_ARG, "null pointer");
    if (n_symbols > 0 && (!x || !mean || !scale || !packed)) return fail(FLIC_E_ARG, "null pointer");
    const EncodeWorkspace w = carve(workspace, n_symbols, n_streams);
    if (workspace_bytes < w.bytes)
        return fail(FLIC_E_CAPACITY, "workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w.bytes);
    FLIC_CUDA(flic::launch_rans_encode(x, mean, scale, stream_offsets, n_streams, init_states, w.scratch,
                                       w.counts, final_states, status, st));
    FLIC_CUDA(flic::launch_scan_counts(w.counts, n_streams, word_offsets, w.scan_tmp, st));
    FLIC_CUDA(flic::launch_pack_words(w.scratch, stream_offsets, word_offsets, n_streams, packed,
                                      packed_capacity, status, st));
    g_launches += 5;
    return 0;
}

int flic_rans_decode_resume(const uint32_t* packed, const int64_t* word_offsets,
                            const uint64_t* states_in, const int64_t* words_left_in, const float* mean,
                            const float* scale, const int64_t* stream_offsets, int64_t n_streams,
                            float* x_out, uint64_t* end_states, int64_t* words_left_out, int32_t* status,
                            int check_end, flic_cuda_stream_t stream) {
    if (n_streams < 0) return fail(FLIC_E_ARG, "n_streams < 0");
    if (n_streams == 0) return 0;
    if (!word_offsets || !states_in || !stream_offsets || !end_states || !status)
        return fail(FLIC_E_ARG, "null pointer");
    const flic::WordsLeft left = {words_left_in, words_left_out};
    FLIC_CUDA(flic::launch_rans_decode(packed, word_offsets, states_in, mean, scale, stream_offsets,
                                       n_streams, x_out, end_states, status, check_end, left, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_rans_decode(const uint32_t* packed, const int64_t* word_offsets,
                     const uint64_t* final_states, const float* mean, const float* scale,
                     const int64_t* stream_offsets, int64_t n_streams, float* x_out,
                     uint64_t* end_states, int32_t* status, int check_end,
                     flic_cuda_stream_t stream) {
    return flic_rans_decode_resume(packed, word_offsets, final_states, nullptr, mean, scale, stream_offsets,
                                   n_streams, x_out, end_states, nullptr, status, check_end, stream);
}

int flic_gather_words(const uint32_t* src, const int64_t* src_offsets, const int64_t* dst_starts,
                      int64_t n_streams, uint32_t* dst, int64_t dst_capacity, int32_t* status,
                      flic_cuda_stream_t stream) {
    if (n_streams < 0 || dst_capacity < 0) return fail(FLIC_E_ARG, "negative size");
    if (n_streams == 0) return 0;
    if (!src_offsets || !dst_starts || (dst_capacity > 0 && (!src || !dst))) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_gather_words(src, src_offsets, dst_starts, n_streams, dst, dst_capacity, status,
                                        (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_couple_add_round(float* x, const float* t, int64_t batch, int64_t channels, int64_t a_ch,
                          int64_t hw, int direction, int nbits, flic_cuda_stream_t stream) {
    if (batch < 0 || channels < 0 || a_ch < 0 || a_ch > channels || hw < 0) return fail(FLIC_E_ARG, "bad shape");
    if (direction != 1 && direction != -1) return fail(FLIC_E_ARG, "direction must be +1 or -1");
    if (nbits < 0 || nbits > 23) return fail(FLIC_E_ARG, "nbits out of range");
    if (batch * (channels - a_ch) * hw == 0) return 0;
    if (!x || !t) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_couple_add_round(x, t, batch, channels, a_ch, hw, (float)direction, nbits,
                                            (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_u8_to_grid(const uint8_t* src, float* dst, int64_t n, flic_cuda_stream_t stream) {
    if (n < 0) return fail(FLIC_E_ARG, "n < 0");
    if (n == 0) return 0;
    if (!src || !dst) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_u8_to_grid(src, dst, n, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_grid_to_u8(const float* src, uint8_t* dst, int64_t n, int32_t* status_word,
                    flic_cuda_stream_t stream) {
    if (n < 0) return fail(FLIC_E_ARG, "n < 0");
    if (n == 0) return 0;
    if (!src || !dst || !status_word) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_grid_to_u8(src, dst, n, status_word, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_permute_channels(const float* src, float* dst, const int32_t* perm, int64_t batch,
                          int64_t channels, int64_t hw, flic_cuda_stream_t stream) {
    if (batch < 0 || channels < 0 || hw < 0) return fail(FLIC_E_ARG, "bad shape");
    if (batch * channels * hw == 0) return 0;
    if (!src || !dst || !perm || src == dst) return fail(FLIC_E_ARG, "null or aliased pointer");
    FLIC_CUDA(flic::launch_permute_channels(src, dst, perm, batch, channels, hw, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_squeeze(const float* src, float* dst, int64_t batch, int64_t C, int64_t H, int64_t W,
                 int scale, int direction, flic_cuda_stream_t stream) {
    if (batch < 0 || C < 0 || H < 0 || W < 0 || scale < 1) return fail(FLIC_E_ARG, "bad shape");
    if (H % scale || W % scale) return fail(FLIC_E_ARG, "H, W must be multiples of scale");
    if (direction != 1 && direction != -1) return fail(FLIC_E_ARG, "direction must be +1 or -1");
    if (batch * C * H * W == 0) return 0;
    if (!src || !dst || src == dst) return fail(FLIC_E_ARG, "null or aliased pointer");
    FLIC_CUDA(flic::launch_squeeze(src, dst, batch, C, H, W, scale, direction, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_dlogistic_log_prob(const float* x, const float* mean, const float* logscale, int64_t batch,
                            int64_t per_item, int nbits, float eps, float* logp_out, float* sum_out,
                            flic_cuda_stream_t stream) {
    if (batch < 0 || per_item < 0 || nbits < 0 || nbits > 23) return fail(FLIC_E_ARG, "bad argument");
    if (batch == 0) return 0;
    if (batch > 0x7fffffffll) return fail(FLIC_E_ARG, "batch too large");
    if ((per_item > 0 && (!x || !mean || !logscale)) || (!logp_out && !sum_out)) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_dlogistic_log_prob(x, mean, logscale, batch, per_item, nbits, eps, logp_out, sum_out,
                                              (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int flic_dlogistic_sample(const float* u, const float* mean, const float* logscale, int64_t n, int nbits,
                          float* out, flic_cuda_stream_t stream) {
    if (n < 0 || nbits < 0 || nbits > 23) return fail(FLIC_E_ARG, "bad argument");
    if (n == 0) return 0;
    if (!u || !mean || !logscale || !out) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(flic::launch_dlogistic_sample(u, mean, logscale, n, nbits, out, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Host entry points
 * ------------------------------------------------------------------------------------------ */

struct flic_codec {
    int device;
    int64_t max_symbols, max_streams;
    // A call is cut into chunks of whole streams (at most slot_symbols symbols / slot_streams
    // streams each) that rotate through kSlots device slots, each with its own CUDA stream: while
    // chunk c is in its kernels, chunk c+1 is on its way up the PCIe link and chunk c-1 on its
    // way down, so both copy engines and the SMs are busy at once.
    static constexpr int kSlots = 3;
    int64_t slot_symbols, slot_streams;
    cudaStream_t streams[kSlots];
    cudaEvent_t done[kSlots];
    struct Slot {
        float *x, *mean, *scale;
        int64_t* offsets;       // this chunk's stream offsets, rebased to start at 0
        uint32_t* packed;
        int64_t* word_offsets;
        uint64_t* states;
        uint64_t* end_states;
        int32_t* status;
        void* workspace;
        int64_t workspace_bytes;
        int64_t* h_word_offsets;  // pinned staging for the chunk-local word offsets
        int64_t* h_offsets;       // pinned staging for the chunk-local stream offsets
        bool busy;                // done[] has been recorded for work that reads the staging arrays
    } slot[kSlots];
};

static void free_slots(flic_codec* c) {
    for (int i = 0; i < flic_codec::kSlots; ++i) {
        flic_codec::Slot& s = c->slot[i];
        cudaFree(s.x); cudaFree(s.mean); cudaFree(s.scale); cudaFree(s.offsets); cudaFree(s.packed);
        cudaFree(s.word_offsets); cudaFree(s.states); cudaFree(s.end_states); cudaFree(s.status);
        cudaFree(s.workspace);
        cudaFreeHost(s.h_word_offsets);
        cudaFreeHost(s.h_offsets);
        memset((void*)&s, 0, sizeof s);
    }
}

// Sizes are recorded only once every slot exists: after a failed (re)allocation the codec owns
// nothing and reports a capacity of zero instead of sizes it does not have.
static cudaError_t alloc_slots(flic_codec* c, int64_t ns, int64_t nt) {
    c->slot_symbols = 0;
    c->slot_streams = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; e == cudaSuccess && i < flic_codec::kSlots; ++i) {
        flic_codec::Slot& s = c->slot[i];
        s.workspace_bytes = flic_encode_workspace_bytes(ns, nt);
        if (e == cudaSuccess) e = cudaMalloc(&s.x, sizeof(float) * ns);
        if (e == cudaSuccess) e = cudaMalloc(&s.mean, sizeof(float) * ns);
        if (e == cudaSuccess) e = cudaMalloc(&s.scale, sizeof(float) * ns);
        if (e == cudaSuccess) e = cudaMalloc(&s.offsets, sizeof(int64_t) * (nt + 1));
        if (e == cudaSuccess) e = cudaMalloc(&s.packed, sizeof(uint32_t) * ns);
        if (e == cudaSuccess) e = cudaMalloc(&s.word_offsets, sizeof(int64_t) * (nt + 1));
        if (e == cudaSuccess) e = cudaMalloc(&s.states, sizeof(uint64_t) * nt);
        if (e == cudaSuccess) e = cudaMalloc(&s.end_states, sizeof(uint64_t) * nt);
        if (e == cudaSuccess) e = cudaMalloc(&s.status, sizeof(int32_t) * nt);
        if (e == cudaSuccess) e = cudaMalloc(&s.workspace, (size_t)s.workspace_bytes);
        if (e == cudaSuccess) e = cudaMallocHost(&s.h_word_offsets, sizeof(int64_t) * (nt + 8));
        if (e == cudaSuccess) e = cudaMallocHost(&s.h_offsets, sizeof(int64_t) * (nt + 8));   // + scalars of the single-stream calls
    }
    if (e != cudaSuccess) {
        free_slots(c);
        return e;
    }
    c->slot_symbols = ns;
    c->slot_streams = nt;
    return e;
}

static void codec_free(flic_codec* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_slots(c);
    for (int i = 0; i < flic_codec::kSlots; ++i) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
    }
    delete c;
}

static cudaError_t sync_all(flic_codec* c) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < flic_codec::kSlots; ++i) {
        const cudaError_t ei = cudaStreamSynchronize(c->streams[i]);
        if (e == cudaSuccess) e = ei;
    }
    return e;
}

// Slots big enough for one stream of `symbols` symbols (a chunk is at least one whole stream).
static int ensure_slot_symbols(flic_codec* c, int64_t symbols) {
    if (symbols <= c->slot_symbols) return 0;
    FLIC_CUDA(sync_all(c));
    const int64_t streams = c->slot_streams > 0 ? c->slot_streams : (c->max_streams < (1 << 20) ? c->max_streams : (1 << 20));
    free_slots(c);
    const cudaError_t e = alloc_slots(c, symbols, streams);
    if (e != cudaSuccess) return cuda_fail(e, "growing codec slots");
    return 0;
}

static int64_t chunk_symbols_target() {
    // 16 Mi symbols per chunk: 192 MB of float inputs (~3.5 ms on a PCIe 5 x16 link) against
    // ~0.3 ms of kernels, and enough streams per launch (>= 2700 at 6144 symbols) to fill the SMs.
    const char* env = getenv("FLIC_CODEC_CHUNK_SYMBOLS");
    if (env) {
        const long long v = atoll(env);
        if (v > 0) return (int64_t)v;
    }
    return (int64_t)16 << 20;
}

int flic_codec_create(int device, int64_t max_symbols, int64_t max_streams, flic_codec** out) {
    if (!out || max_symbols < 1 || max_streams < 1) return fail(FLIC_E_ARG, "bad codec size");
    flic_codec* c = new (std::nothrow) flic_codec();
    if (!c) return fail(FLIC_E_NOMEM, "out of host memory");
    memset((void*)c, 0, sizeof *c);
    c->device = device;
    c->max_symbols = max_symbols;
    c->max_streams = max_streams;
    const int64_t target = chunk_symbols_target();
    cudaError_t e = cudaSetDevice(device);
    for (int i = 0; e == cudaSuccess && i < flic_codec::kSlots; ++i) {
        e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess)
        e = alloc_slots(c, max_symbols < target ? max_symbols : target, max_streams < (1 << 20) ? max_streams : (1 << 20));
    if (e != cudaSuccess) {
        codec_free(c);
        return cuda_fail(e, "flic_codec_create");
    }
    *out = c;
    return 0;
}

void flic_codec_destroy(flic_codec* codec) { codec_free(codec); }

int flic_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(FLIC_E_ARG, "bad host alloc");
    FLIC_CUDA(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return 0;
}

void flic_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

static int check_offsets(const int64_t* off, int64_t n_streams) {
    if (off[0] != 0) return fail(FLIC_E_ARG, "stream_offsets[0] != 0");
    for (int64_t s = 0; s < n_streams; ++s)
        if (off[s + 1] < off[s]) return fail(FLIC_E_ARG, "stream_offsets not monotone at %lld", (long long)s);
    return 0;
}

// Last stream (exclusive) of the chunk that starts at s0: as many whole streams as fit `cap`
// symbols (at most a slot).  Returns s0 when the first stream alone is larger than a slot.
static int64_t chunk_end(const flic_codec* c, const int64_t* off, int64_t n_streams, int64_t s0, int64_t cap) {
    int64_t lo = s0, hi = n_streams < s0 + c->slot_streams ? n_streams : s0 + c->slot_streams;
    const int64_t limit = off[s0] + cap;
    while (lo < hi) {  // largest s1 in (s0, hi] with off[s1] <= limit
        const int64_t mid = lo + (hi - lo + 1) / 2;
        if (off[mid] <= limit) lo = mid; else hi = mid - 1;
    }
    if (lo == s0 && s0 < n_streams && off[s0 + 1] - off[s0] <= c->slot_symbols) lo = s0 + 1;   // one stream above the cap still fits a slot
    return lo;
}

// Symbols to aim for in the chunk starting at symbol `done` of `total`: the pipeline's fill (nothing
// to overlap the first upload with) and drain (nothing overlaps the last kernels and download)
// cost one chunk each, so the first chunk is an eighth of a slot and the last ones halve.
static int64_t chunk_cap(const flic_codec* c, int64_t chunk, int64_t done, int64_t total) {
    const int64_t small = c->slot_symbols / 8 > 0 ? c->slot_symbols / 8 : 1;
    if (chunk == 0) return small;
    int64_t cap = (total - done) / 2;
    if (cap < small) cap = small;
    return cap < c->slot_symbols ? cap : c->slot_symbols;
}

static int codec_encode_impl(flic_codec* c, const float* x, const float* mean, const float* scale,
                             const int64_t* stream_offsets, int64_t n_streams, uint32_t* words_out,
                             int64_t words_capacity, int64_t* word_offsets_out, uint64_t* states_out,
                             int32_t* status_out, int64_t* n_words_out) {
    if (!c || n_streams < 0 || !stream_offsets || !word_offsets_out) return fail(FLIC_E_ARG, "bad argument");
    if (int rc = check_offsets(stream_offsets, n_streams)) return rc;
    const int64_t n_symbols = stream_offsets[n_streams];
    if (n_symbols > c->max_symbols || n_streams > c->max_streams)
        return fail(FLIC_E_CAPACITY, "codec sized for %lld symbols / %lld streams", (long long)c->max_symbols,
                    (long long)c->max_streams);
    word_offsets_out[0] = 0;
    if (n_words_out) *n_words_out = 0;
    if (n_streams == 0) return 0;
    if (!states_out || !status_out || (n_symbols > 0 && (!x || !mean || !scale || !words_out)))
        return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(cudaSetDevice(c->device));

    const int64_t* off = stream_offsets;
    int64_t total_words = 0;       // words of all finalised chunks
    bool overflow = false;
    struct Pending { int slot; int64_t s0, s1; bool active; } prev = {0, 0, 0, false};

    // Second half of a chunk: its word counts are on the host now, so the words can be copied to
    // their final place and the caller's word offsets filled in.
    auto finalize = [&](const Pending& p) -> int {
        flic_codec::Slot& sl = c->slot[p.slot];
        FLIC_CUDA(cudaEventSynchronize(c->done[p.slot]));
        const int64_t ns = p.s1 - p.s0;
        const int64_t* hw = sl.h_word_offsets;
        const int64_t chunk_words = hw[ns];
        for (int64_t i = 1; i <= ns; ++i) word_offsets_out[p.s0 + i] = total_words + hw[i];
        if (total_words + chunk_words > words_capacity) overflow = true;
        else if (chunk_words > 0)
            FLIC_CUDA(cudaMemcpyAsync(words_out + total_words, sl.packed, sizeof(uint32_t) * chunk_words,
                                      cudaMemcpyDeviceToHost, c->streams[p.slot]));
        total_words += chunk_words;
        return 0;
    };

    int64_t s0 = 0;
    for (int64_t chunk = 0; s0 < n_streams; ++chunk) {
        int64_t s1 = chunk_end(c, off, n_streams, s0, chunk_cap(c, chunk, off[s0], n_symbols));
        if (s1 == s0) {  // one stream larger than a slot: finish what is in flight, grow, retry
            if (prev.active) { if (int rc = finalize(prev)) return rc; prev.active = false; }
            if (int rc = ensure_slot_symbols(c, off[s0 + 1] - off[s0])) return rc;
            s1 = chunk_end(c, off, n_streams, s0, c->slot_symbols);
        }
        const int slot = (int)(chunk % flic_codec::kSlots);
        flic_codec::Slot& sl = c->slot[slot];
        cudaStream_t st = c->streams[slot];
        const int64_t a = off[s0], n = off[s1] - a, ns = s1 - s0;
        // The slot's staging arrays were last read by the chunk that used this slot kSlots chunks
        // ago; its `done` event was waited for in finalize(), so they are free.  Offsets are rebased
        // to the chunk so that the kernels index the slot's arrays from 0.
        for (int64_t i = 0; i <= ns; ++i) sl.h_offsets[i] = off[s0 + i] - a;
        if (n > 0) {
            FLIC_CUDA(cudaMemcpyAsync(sl.x, x + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
            FLIC_CUDA(cudaMemcpyAsync(sl.mean, mean + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
            FLIC_CUDA(cudaMemcpyAsync(sl.scale, scale + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        }
        FLIC_CUDA(cudaMemcpyAsync(sl.offsets, sl.h_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice, st));
        const EncodeWorkspace w = carve(sl.workspace, c->slot_symbols, c->slot_streams);
        FLIC_CUDA(flic::launch_rans_encode(sl.x, sl.mean, sl.scale, sl.offsets, ns, nullptr, w.scratch,
                                           w.counts, sl.states, sl.status, st));
        FLIC_CUDA(flic::launch_scan_counts(w.counts, ns, sl.word_offsets, w.scan_tmp, st));
        FLIC_CUDA(flic::launch_pack_words(w.scratch, sl.offsets, sl.word_offsets, ns, sl.packed, c->slot_symbols,
                                          sl.status, st));
        g_launches += 5;
        FLIC_CUDA(cudaMemcpyAsync(sl.h_word_offsets, sl.word_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyDeviceToHost, st));
        FLIC_CUDA(cudaMemcpyAsync(states_out + s0, sl.states, sizeof(uint64_t) * ns, cudaMemcpyDeviceToHost, st));
        FLIC_CUDA(cudaMemcpyAsync(status_out + s0, sl.status, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
        FLIC_CUDA(cudaEventRecord(c->done[slot], st));
        if (prev.active) { if (int rc = finalize(prev)) return rc; }
        prev = {slot, s0, s1, true};
        s0 = s1;
    }
    if (prev.active) { if (int rc = finalize(prev)) return rc; }
    FLIC_CUDA(sync_all(c));
    if (n_words_out) *n_words_out = total_words;
    if (overflow)
        return fail(FLIC_E_CAPACITY, "words_out holds %lld words, %lld needed", (long long)words_capacity, (long long)total_words);
    return 0;
}

// Whatever the outcome, nothing of the codec's is in flight when a host entry point returns: an
// early return on an error would otherwise leave asynchronous copies writing into buffers the
// caller is about to free.
int flic_codec_encode(flic_codec* c, const float* x, const float* mean, const float* scale,
                      const int64_t* stream_offsets, int64_t n_streams, uint32_t* words_out,
                      int64_t words_capacity, int64_t* word_offsets_out, uint64_t* states_out,
                      int32_t* status_out, int64_t* n_words_out) {
    const int rc = codec_encode_impl(c, x, mean, scale, stream_offsets, n_streams, words_out, words_capacity,
                                     word_offsets_out, states_out, status_out, n_words_out);
    if (rc != 0 && c) sync_all(c);
    return rc;
}

static int codec_decode_impl(flic_codec* c, const uint32_t* words, const int64_t* word_offsets,
                             const uint64_t* states, const float* mean, const float* scale,
                             const int64_t* stream_offsets, int64_t n_streams, float* x_out,
                             uint64_t* end_states_out, int32_t* status_out) {
    if (!c || n_streams < 0 || !stream_offsets || !word_offsets) return fail(FLIC_E_ARG, "bad argument");
    if (n_streams == 0) return 0;
    if (int rc = check_offsets(stream_offsets, n_streams)) return rc;
    if (int rc = check_offsets(word_offsets, n_streams)) return rc;
    const int64_t n_symbols = stream_offsets[n_streams];
    const int64_t n_words = word_offsets[n_streams];
    if (n_symbols > c->max_symbols || n_streams > c->max_streams || n_words > c->max_symbols)
        return fail(FLIC_E_CAPACITY, "codec sized for %lld symbols / %lld streams", (long long)c->max_symbols,
                    (long long)c->max_streams);
    if (!states || !status_out || (n_symbols > 0 && (!mean || !scale || !x_out)) || (n_words > 0 && !words))
        return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(cudaSetDevice(c->device));
    const int64_t* off = stream_offsets;
    int64_t s0 = 0;
    for (int64_t chunk = 0; s0 < n_streams; ++chunk) {
        int64_t s1 = chunk_end(c, off, n_streams, s0, chunk_cap(c, chunk, off[s0], n_symbols));
        if (s1 == s0) {
            if (int rc = ensure_slot_symbols(c, off[s0 + 1] - off[s0])) return rc;
            s1 = chunk_end(c, off, n_streams, s0, c->slot_symbols);
        }
        const int64_t a = off[s0], n = off[s1] - a, ns = s1 - s0;
        const int64_t wa = word_offsets[s0], nw = word_offsets[s1] - wa;
        if (nw > c->slot_symbols) {  // a valid stream never has more words than symbols
            if (int rc = ensure_slot_symbols(c, nw)) return rc;
        }
        const int slot = (int)(chunk % flic_codec::kSlots);
        flic_codec::Slot& sl = c->slot[slot];
        cudaStream_t st = c->streams[slot];
        // offsets are rebased to the chunk on the host; the staging arrays are free once the work
        // that last read them (this slot, kSlots chunks ago) has finished
        if (sl.busy) FLIC_CUDA(cudaEventSynchronize(c->done[slot]));
        for (int64_t i = 0; i <= ns; ++i) {
            sl.h_offsets[i] = off[s0 + i] - a;
            sl.h_word_offsets[i] = word_offsets[s0 + i] - wa;
        }
        if (nw > 0) FLIC_CUDA(cudaMemcpyAsync(sl.packed, words + wa, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, st));
        FLIC_CUDA(cudaMemcpyAsync(sl.word_offsets, sl.h_word_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice, st));
        FLIC_CUDA(cudaMemcpyAsync(sl.states, states + s0, sizeof(uint64_t) * ns, cudaMemcpyHostToDevice, st));
        if (n > 0) {
            FLIC_CUDA(cudaMemcpyAsync(sl.mean, mean + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
            FLIC_CUDA(cudaMemcpyAsync(sl.scale, scale + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        }
        FLIC_CUDA(cudaMemcpyAsync(sl.offsets, sl.h_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice, st));
        FLIC_CUDA(cudaEventRecord(c->done[slot], st));
        sl.busy = true;
        FLIC_CUDA(flic::launch_rans_decode(sl.packed, sl.word_offsets, sl.states, sl.mean, sl.scale,
                                           sl.offsets, ns, sl.x, sl.end_states, sl.status, 1, flic::WordsLeft{nullptr, nullptr}, st));
        g_launches += 1;
        if (n > 0) FLIC_CUDA(cudaMemcpyAsync(x_out + a, sl.x, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        if (end_states_out)
            FLIC_CUDA(cudaMemcpyAsync(end_states_out + s0, sl.end_states, sizeof(uint64_t) * ns, cudaMemcpyDeviceToHost, st));
        FLIC_CUDA(cudaMemcpyAsync(status_out + s0, sl.status, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
        s0 = s1;
    }
    FLIC_CUDA(sync_all(c));
    return 0;
}

int flic_codec_decode(flic_codec* c, const uint32_t* words, const int64_t* word_offsets,
                      const uint64_t* states, const float* mean, const float* scale,
                      const int64_t* stream_offsets, int64_t n_streams, float* x_out,
                      uint64_t* end_states_out, int32_t* status_out) {
    const int rc = codec_decode_impl(c, words, word_offsets, states, mean, scale, stream_offsets, n_streams, x_out,
                                     end_states_out, status_out);
    if (rc != 0 && c) sync_all(c);
    return rc;
}

// The same copies as flic_codec_encode followed by flic_codec_decode -- same chunks, same three
// streams, same host buffers -- with no kernel in between: what the host side of this box can
// move for this call pattern, i.e. the ceiling the end-to-end number is measured against.
// words / n_words stand for the compressed payload (copied down chunk by chunk in the encode
// leg and up again in the decode leg, each chunk taking its share by symbol count).
int flic_codec_probe_copies(flic_codec* c, const float* x, const float* mean, const float* scale,
                            const int64_t* stream_offsets, int64_t n_streams, uint32_t* words,
                            int64_t n_words, float* x_out) {
    if (!c || n_streams < 0 || !stream_offsets) return fail(FLIC_E_ARG, "bad argument");
    if (n_streams == 0) return 0;
    if (int rc = check_offsets(stream_offsets, n_streams)) return rc;
    const int64_t n_symbols = stream_offsets[n_streams];
    if (n_symbols > c->max_symbols || n_streams > c->max_streams || n_words > n_symbols)
        return fail(FLIC_E_CAPACITY, "codec sized for %lld symbols / %lld streams", (long long)c->max_symbols,
                    (long long)c->max_streams);
    if (n_symbols > 0 && (!x || !mean || !scale || !x_out || (n_words > 0 && !words))) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(cudaSetDevice(c->device));
    const int64_t* off = stream_offsets;
    auto leg = [&](bool decode) -> int {
        int64_t s0 = 0, wdone = 0;
        for (int64_t chunk = 0; s0 < n_streams; ++chunk) {
            int64_t s1 = chunk_end(c, off, n_streams, s0, chunk_cap(c, chunk, off[s0], n_symbols));
            if (s1 == s0) {
                if (int rc = ensure_slot_symbols(c, off[s0 + 1] - off[s0])) return rc;
                s1 = chunk_end(c, off, n_streams, s0, c->slot_symbols);
            }
            flic_codec::Slot& sl = c->slot[chunk % flic_codec::kSlots];
            cudaStream_t st = c->streams[chunk % flic_codec::kSlots];
            const int64_t a = off[s0], n = off[s1] - a, ns = s1 - s0;
            const int64_t wend = n_symbols > 0 ? (int64_t)((double)n_words * (double)off[s1] / (double)n_symbols) : 0;
            const int64_t nw = wend - wdone;
            if (!decode) {
                if (n > 0) {
                    FLIC_CUDA(cudaMemcpyAsync(sl.x, x + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
                    FLIC_CUDA(cudaMemcpyAsync(sl.mean, mean + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
                    FLIC_CUDA(cudaMemcpyAsync(sl.scale, scale + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
                }
                FLIC_CUDA(cudaMemcpyAsync(sl.offsets, sl.h_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice, st));
                FLIC_CUDA(cudaMemcpyAsync(sl.h_word_offsets, sl.word_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyDeviceToHost, st));
                if (nw > 0) FLIC_CUDA(cudaMemcpyAsync(words + wdone, sl.packed, sizeof(uint32_t) * nw, cudaMemcpyDeviceToHost, st));
            } else {
                if (nw > 0) FLIC_CUDA(cudaMemcpyAsync(sl.packed, words + wdone, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, st));
                FLIC_CUDA(cudaMemcpyAsync(sl.word_offsets, sl.h_word_offsets, sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice, st));
                if (n > 0) {
                    FLIC_CUDA(cudaMemcpyAsync(sl.mean, mean + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
                    FLIC_CUDA(cudaMemcpyAsync(sl.scale, scale + a, sizeof(float) * n, cudaMemcpyHostToDevice, st));
                    FLIC_CUDA(cudaMemcpyAsync(x_out + a, sl.x, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
                }
            }
            wdone = wend;
            s0 = s1;
        }
        FLIC_CUDA(sync_all(c));
        return 0;
    };
    int rc = leg(false);
    if (rc == 0) rc = leg(true);
    if (rc != 0) sync_all(c);
    return rc;
}

// One stream, the reference's own signatures.  Inputs go up from the caller's (usually pageable)
// arrays, the small outputs and the worst-case word buffer come down in the same batch, and the
// call synchronises once: scalars that have to outlive the enqueue sit in the slot's pinned
// staging arrays, not on this stack frame.
int flic_rans_encode_single(flic_codec* c, uint64_t state, int64_t n, const float* x,
                            const float* mean, const float* scale, uint32_t* buffer_out,
                            int64_t* n_words_out, uint64_t* state_out, int32_t* status_out) {
    if (!c || n < 0 || !n_words_out || !state_out) return fail(FLIC_E_ARG, "bad argument");
    if (n > c->max_symbols) return fail(FLIC_E_CAPACITY, "codec sized for %lld symbols", (long long)c->max_symbols);
    *n_words_out = 0;
    *state_out = state;
    if (status_out) *status_out = 0;
    if (n == 0) return 0;
    if (!x || !mean || !scale || !buffer_out) return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(cudaSetDevice(c->device));
    if (int rc = ensure_slot_symbols(c, n)) return rc;
    flic_codec::Slot& s = c->slot[0];
    cudaStream_t st = c->streams[0];
    if (s.busy) FLIC_CUDA(cudaEventSynchronize(c->done[0]));
    s.busy = false;
    // pinned staging: h_offsets = {0, n, state, status}, h_word_offsets = {0, n_words}
    s.h_offsets[0] = 0;
    s.h_offsets[1] = n;
    s.h_offsets[2] = (int64_t)state;
    int rc = 0;
    cudaError_t e = cudaMemcpyAsync(s.x, x, sizeof(float) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.mean, mean, sizeof(float) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.scale, scale, sizeof(float) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.offsets, s.h_offsets, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.end_states, s.h_offsets + 2, sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        rc = flic_rans_encode(s.x, s.mean, s.scale, s.offsets, 1, n, s.end_states, s.workspace, s.workspace_bytes,
                              s.packed, n, s.word_offsets, s.states, s.status, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(s.h_word_offsets, s.word_offsets, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(s.h_offsets + 2, s.states, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(s.h_offsets + 3, s.status, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    // a stream emits at most one word per symbol: the whole worst-case buffer comes down with the rest
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(buffer_out, s.packed, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return cuda_fail(e, "flic_rans_encode_single");
    if (rc) return rc;
    int32_t status;
    memcpy(&status, s.h_offsets + 3, sizeof status);
    *n_words_out = s.h_word_offsets[1];
    *state_out = (uint64_t)s.h_offsets[2];
    if (status_out) *status_out = status;
    if (status) return fail(FLIC_E_STATUS, "stream status 0x%x", status);
    return 0;
}

int flic_rans_decode_single(flic_codec* c, uint64_t state, const uint32_t* buffer_reversed,
                            int64_t n_buffer, int64_t n, const float* mean_reversed,
                            const float* scale_reversed, float* message_out, uint64_t* state_out,
                            int32_t* status_out) {
    if (!c || n < 0 || n_buffer < 0 || !state_out) return fail(FLIC_E_ARG, "bad argument");
    if (n > c->max_symbols || n_buffer > c->max_symbols)
        return fail(FLIC_E_CAPACITY, "codec sized for %lld symbols", (long long)c->max_symbols);
    *state_out = state;
    if (status_out) *status_out = 0;
    if (n == 0) return 0;
    if (!mean_reversed || !scale_reversed || !message_out || (n_buffer > 0 && !buffer_reversed))
        return fail(FLIC_E_ARG, "null pointer");
    FLIC_CUDA(cudaSetDevice(c->device));
    if (int rc = ensure_slot_symbols(c, n > n_buffer ? n : n_buffer)) return rc;
    flic_codec::Slot& s = c->slot[0];
    cudaStream_t st = c->streams[0];
    if (s.busy) FLIC_CUDA(cudaEventSynchronize(c->done[0]));
    s.busy = false;
    // The caller's reversal (trainer.py:317) is undone on the device: the arrays go up as they are
    // and three small kernels reverse them (the decode kernels walk a stream from its end).
    s.h_offsets[0] = 0;
    s.h_offsets[1] = n;
    s.h_offsets[2] = (int64_t)state;
    s.h_word_offsets[0] = 0;
    s.h_word_offsets[1] = n_buffer;
    int rc = 0;
    float* const up = (float*)s.workspace;                 // n floats of upload space (the encode scratch is idle)
    cudaError_t e = cudaSuccess;
    if (n_buffer > 0) {
        e = cudaMemcpyAsync(up, buffer_reversed, sizeof(uint32_t) * n_buffer, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = flic::launch_reverse_u32((const uint32_t*)up, s.packed, n_buffer, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.x, mean_reversed, sizeof(float) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = flic::launch_reverse_u32((const uint32_t*)s.x, (uint32_t*)s.mean, n, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.x, scale_reversed, sizeof(float) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = flic::launch_reverse_u32((const uint32_t*)s.x, (uint32_t*)s.scale, n, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.offsets, s.h_offsets, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.word_offsets, s.h_word_offsets, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.states, s.h_offsets + 2, sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        rc = flic_rans_decode(s.packed, s.word_offsets, s.states, s.mean, s.scale, s.offsets, 1, s.x, s.end_states,
                              s.status, 0, st);
    // the message goes back reversed, as the reference returns it
    if (e == cudaSuccess && rc == 0) e = flic::launch_reverse_u32((const uint32_t*)s.x, (uint32_t*)up, n, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(message_out, up, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(s.h_offsets + 2, s.end_states, sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(s.h_offsets + 3, s.status, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return cuda_fail(e, "flic_rans_decode_single");
    if (rc) return rc;
    g_launches += 4;
    int32_t status;
    memcpy(&status, s.h_offsets + 3, sizeof status);
    *state_out = (uint64_t)s.h_offsets[2];
    if (status_out) *status_out = status;
    if (status) return fail(FLIC_E_STATUS, "stream status 0x%x", status);
    return 0;
}

}  // extern "C"
