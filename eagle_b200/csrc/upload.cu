// Host frames -> device memory for the reference-facing call (the reference hands get_coordinates a Python list of
// pageable numpy frames, eagle/models/coordinate_model.py:188,221).
//
// A pageable buffer cannot be DMA'd: it is either copied into page-locked staging first or the driver does that itself
// on one thread (8-10 GB/s measured).  Staging a whole chunk in a large pinned buffer makes every byte cross host DRAM
// three times (read the frame, write the staging buffer, DMA reads it back): on the bench boxes that is the limit
// (about 100 GB/s of DRAM traffic = 34 GB/s of frames with both legs running) long before PCIe Gen5 (55 GB/s).  Here each
// worker thread copies 4 MiB slices into a small page-locked ring of its own and issues the H2D of every slice at once on
// its own stream, so copy and DMA interleave at slice granularity: 44-47 GB/s from pageable frames with 8-16 threads on
// the same boxes (tools/upload_probe.py), 39 GB/s with 4.
#include <atomic>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#ifdef EGL_BENCH_VARIANTS
#include <immintrin.h>
#endif

namespace egl {

constexpr int kUpMaxThreads = 32;
constexpr int kUpSlots = 3;
static size_t kUpSlice = 4u << 20;  // 256 KiB ... 4 MiB swept on B200 boxes: per-slice overhead loses below 2 MiB

struct UploadWorker {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[kUpSlots] = {};
    unsigned char* slot[kUpSlots] = {};
    int device = -1;
};

struct UploadJob {
    UploadWorker* w;
    const void* const* frames;
    std::atomic<long long>* next;  // slices are handed out one at a time, so no thread is left with a longer tail
    long long total;
    int per_frame;                 // slices per frame
    size_t bytes;
    unsigned char* dst;
    int device;
    cudaError_t err;
};

#ifdef EGL_BENCH_VARIANTS
// A/B of the staging copy (tools/upload_variants_probe.py): 0 memcpy (the shipped choice), 1 AVX2 loads + non-temporal
// stores (no read-for-ownership of the ring, nothing of it left in the caches), 2 `rep movsb`
static int g_up_copy_mode = 0;
static unsigned g_up_alloc_flags = cudaHostAllocDefault;
__attribute__((target("avx2"))) static void copy_nontemporal(unsigned char* d, const unsigned char* s, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i)), b = _mm256_loadu_si256((const __m256i*)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + i + 64)), e = _mm256_loadu_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    if (i < n) memcpy(d + i, s + i, n - i);
    _mm_sfence();   // the DMA engine must see the data: non-temporal stores are weakly ordered
}
static void stage_copy(unsigned char* d, const unsigned char* s, size_t n) {
    if (g_up_copy_mode == 1) copy_nontemporal(d, s, n);
    else if (g_up_copy_mode == 2) { __asm__ volatile("rep movsb" : "+D"(d), "+S"(s), "+c"(n) : : "memory"); }
    else memcpy(d, s, n);
}
#else
static inline void stage_copy(unsigned char* d, const unsigned char* s, size_t n) { memcpy(d, s, n); }
constexpr unsigned g_up_alloc_flags = cudaHostAllocDefault;
#endif

// Two independent sets of workers (rings, streams): a second call may run while the first one drains its last DMAs,
// which is how the streaming layer keeps the link busy across chunk boundaries.  A third concurrent call waits.
constexpr int kUpSets = 2;
static UploadWorker g_up_sets[kUpSets][kUpMaxThreads];
static pthread_mutex_t g_up_lock[kUpSets] = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER};

static cudaError_t upload_worker_init(UploadWorker& w, int device) {
    if (w.device == device) return cudaSuccess;
    if (w.device >= 0) {  // the library was last used on another device: rebuild on this one
        cudaStreamDestroy(w.stream);
        for (int s = 0; s < kUpSlots; ++s) { cudaEventDestroy(w.ev[s]); cudaFreeHost(w.slot[s]); }
        w = UploadWorker();
    }
    cudaError_t e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
    for (int s = 0; s < kUpSlots && e == cudaSuccess; ++s) {
        e = cudaEventCreateWithFlags(&w.ev[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&w.slot[s], kUpSlice, g_up_alloc_flags);
    }
    if (e == cudaSuccess) w.device = device;
    return e;
}

static void* upload_thread(void* arg) {
    UploadJob& j = *(UploadJob*)arg;
    j.err = cudaSetDevice(j.device);
    if (j.err != cudaSuccess) return nullptr;
    UploadWorker& w = *j.w;
    int q = 0;
    bool used[kUpSlots] = {};
    for (;;) {
        const long long t = j.next->fetch_add(1, std::memory_order_relaxed);
        if (t >= j.total || j.err != cudaSuccess) break;
        const long long f = t / j.per_frame;
        const size_t o = (size_t)(t % j.per_frame) * kUpSlice;
        const size_t n = j.bytes - o < kUpSlice ? j.bytes - o : kUpSlice;
        const unsigned char* src = (const unsigned char*)j.frames[f] + o;
        unsigned char* dst = j.dst + (size_t)f * j.bytes + o;
        if (used[q]) j.err = cudaEventSynchronize(w.ev[q]);  // the DMA that last read this slot has finished
        if (j.err != cudaSuccess) break;
        stage_copy(w.slot[q], src, n);
        j.err = cudaMemcpyAsync(dst, w.slot[q], n, cudaMemcpyHostToDevice, w.stream);
        if (j.err == cudaSuccess) j.err = cudaEventRecord(w.ev[q], w.stream);
        used[q] = true;
        q = (q + 1) % kUpSlots;
    }
    const cudaError_t e = cudaStreamSynchronize(w.stream);
    if (j.err == cudaSuccess) j.err = e;
    return nullptr;
}

}  // namespace egl

using namespace egl;

extern "C" int egl_upload_frames(const void* const* frames, int n_frames, size_t bytes_per_frame, void* dst, int n_threads) {
    if (n_frames == 0) return 0;
    EGL_REQUIRE(frames && dst, EGL_ERR_NULL, "egl_upload_frames: null pointer");
    EGL_REQUIRE(n_frames > 0 && bytes_per_frame > 0, EGL_ERR_SHAPE, "egl_upload_frames: need n_frames >= 0 and bytes_per_frame > 0");
    for (int i = 0; i < n_frames; ++i) EGL_REQUIRE(frames[i], EGL_ERR_NULL, "egl_upload_frames: frame %d is null", i);
#ifdef EGL_BENCH_VARIANTS
    static const char* slice_env = getenv("EGL_UPLOAD_SLICE_KB");
    if (slice_env && g_up_sets[0][0].device < 0 && g_up_sets[1][0].device < 0) kUpSlice = (size_t)atoi(slice_env) << 10;
    static const char* copy_env = getenv("EGL_UPLOAD_COPY");
    if (copy_env) g_up_copy_mode = atoi(copy_env);
    static const char* wc_env = getenv("EGL_UPLOAD_WC");
    if (wc_env && atoi(wc_env) && g_up_sets[0][0].device < 0 && g_up_sets[1][0].device < 0) g_up_alloc_flags = cudaHostAllocWriteCombined;
#endif
    int nt = n_threads < 1 ? 1 : (n_threads > kUpMaxThreads ? kUpMaxThreads : n_threads);
    const int per_frame = (int)((bytes_per_frame + kUpSlice - 1) / kUpSlice);
    const long long total = (long long)n_frames * per_frame;
    if (nt > total) nt = (int)total;
    std::atomic<long long> next(0);
    int device = 0;
    int rc = cuda_status(cudaGetDevice(&device), "egl_upload_frames: cudaGetDevice");
    if (rc) return rc;
    int set = 0;  // a set of workers serves one call at a time: its rings and streams are shared state
    if (pthread_mutex_trylock(&g_up_lock[0]) != 0) {
        set = 1;
        pthread_mutex_lock(&g_up_lock[1]);
    }
    UploadWorker* g_up = g_up_sets[set];
    UploadJob jobs[kUpMaxThreads];
    pthread_t th[kUpMaxThreads];
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < nt && err == cudaSuccess; ++k) err = upload_worker_init(g_up[k], device);
    int started = 0;
    if (err == cudaSuccess) {
        for (int k = 0; k < nt; ++k) {
            jobs[k] = UploadJob{&g_up[k], frames, &next, total, per_frame, bytes_per_frame, (unsigned char*)dst, device, cudaSuccess};
            if (pthread_create(&th[k], nullptr, upload_thread, &jobs[k]) != 0) break;
            ++started;
        }
        for (int k = 0; k < started; ++k) pthread_join(th[k], nullptr);
        for (int k = 0; k < started; ++k)
            if (jobs[k].err != cudaSuccess) err = jobs[k].err;
        if (started == 0 && err == cudaSuccess) err = cudaErrorLaunchFailure;  // no worker thread could be created (fewer than asked for is fine: the slices are shared out dynamically)
    }
    pthread_mutex_unlock(&g_up_lock[set]);
    return cuda_status(err, "egl_upload_frames");
}
