// upload.cu -- host -> device staging for the tile loader (PD:1828-1894 `_load_bit_data` feeds the GPU).
//
// A datastore hands the loader plain NumPy arrays, i.e. PAGEABLE memory.  cudaMemcpy from pageable
// memory runs at ~11 GB/s on the B200 hosts (one driver thread copying through one small bounce
// buffer): 1.2 s for a 13.4 GB tile whose decode takes 2.8 ms.  m3d_upload stages through a ring of
// pinned slots owned by the context: a few host threads memcpy chunks into free slots while earlier
// slots drain over PCIe with cudaMemcpyAsync, which reaches ~49 GB/s (88 % of the 55.5 GB/s pinned
// rate).  Sources that are already page-locked are copied directly.
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace {

constexpr size_t SLOT_BYTES = (size_t)32 << 20;
constexpr int N_SLOTS = 8;
constexpr int MAX_WORKERS = 8;

struct UploadRing {
    void* slot[N_SLOTS] = {nullptr};
    cudaEvent_t drained[N_SLOTS] = {nullptr};
    bool used[N_SLOTS] = {false};  // drained[s] has been recorded at least once
    bool ready = false;
};

std::mutex g_ring_mu;
std::vector<std::pair<m3d_ctx*, UploadRing*>> g_rings;  // rings live as long as their context

UploadRing* ring_of(m3d_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_ring_mu);
    for (auto& p : g_rings)
        if (p.first == ctx) return p.second;
    UploadRing* r = new UploadRing();
    g_rings.push_back({ctx, r});
    return r;
}

}  // namespace

void m3d_release_upload_ring(m3d_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_ring_mu);
    for (size_t i = 0; i < g_rings.size(); ++i) {
        if (g_rings[i].first != ctx) continue;
        UploadRing* r = g_rings[i].second;
        for (int s = 0; s < N_SLOTS; ++s) {
            if (r->used[s]) cudaEventSynchronize(r->drained[s]);
            if (r->slot[s]) cudaFreeHost(r->slot[s]);
            if (r->drained[s]) cudaEventDestroy(r->drained[s]);
        }
        delete r;
        g_rings.erase(g_rings.begin() + i);
        return;
    }
}

extern "C" int m3d_upload_batch_cb(m3d_ctx* ctx, int n_pieces, const void* const* src_host, void* const* dst_dev,
                                   const int64_t* n_bytes, void* stream, m3d_piece_callback on_piece, void* user) {
    if (!ctx || n_pieces < 0 || (n_pieces > 0 && (!src_host || !dst_dev || !n_bytes)))
        return m3d_fail(M3D_ERR_ARG, "m3d_upload: bad argument");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    struct Chunk {
        const char* src;
        char* dst;
        size_t n;
        bool direct;  // page-locked source queued behind staged chunks: plain DMA, no slot
    };
    std::vector<Chunk> chunks;
    std::vector<int> last_chunk_of;  // last_chunk_of[j] = piece whose final chunk is j, else -1
    for (int p = 0; p < n_pieces; ++p) {
        if (n_bytes[p] < 0) return m3d_fail(M3D_ERR_ARG, "m3d_upload: negative size");
        if (n_bytes[p] == 0) {
            if (on_piece) on_piece(p, user);
            continue;
        }
        if (!src_host[p] || !dst_dev[p]) return m3d_fail(M3D_ERR_ARG, "m3d_upload: null pointer");
        cudaPointerAttributes attr;
        const cudaError_t pe = cudaPointerGetAttributes(&attr, src_host[p]);
        if (pe != cudaSuccess) cudaGetLastError();  // old drivers report unregistered memory as an error
        const bool pinned =
            (pe == cudaSuccess) && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        const size_t total = (size_t)n_bytes[p];
        if (pinned || total <= ((size_t)1 << 20)) {  // page-locked already (or tiny): straight DMA
            // pieces complete on the stream in issue order: staged chunks of earlier pieces go first
            if (chunks.empty()) {
                M3D_CUDA(cudaMemcpyAsync(dst_dev[p], src_host[p], total, cudaMemcpyHostToDevice, st));
                if (on_piece) on_piece(p, user);
                continue;
            }
            chunks.push_back({reinterpret_cast<const char*>(src_host[p]), reinterpret_cast<char*>(dst_dev[p]), total, true});
            last_chunk_of.push_back(p);
            continue;
        }
        for (size_t off = 0; off < total; off += SLOT_BYTES) {
            chunks.push_back({reinterpret_cast<const char*>(src_host[p]) + off, reinterpret_cast<char*>(dst_dev[p]) + off,
                              total - off < SLOT_BYTES ? total - off : SLOT_BYTES, false});
            last_chunk_of.push_back(off + SLOT_BYTES >= total ? p : -1);
        }
    }
    const size_t n_chunks = chunks.size();
    if (n_chunks == 0) return M3D_OK;
    UploadRing* R = ring_of(ctx);
    if (!R->ready) {
        for (int s = 0; s < N_SLOTS; ++s) {
            M3D_CUDA(cudaHostAlloc(&R->slot[s], SLOT_BYTES, cudaHostAllocDefault));
            M3D_CUDA(cudaEventCreateWithFlags(&R->drained[s], cudaEventDisableTiming));
        }
        R->ready = true;
    }
    // chunk j uses slot j % N_SLOTS.  staged[j]: a worker has filled the slot; issued[j]: this thread has
    // enqueued its cudaMemcpyAsync and recorded drained[slot] behind it.  Slots still draining from the
    // previous call are waited for through the same events (R->used).
    std::vector<char> staged(n_chunks, 0), issued(n_chunks, 0);
    std::vector<int> slot_of(n_chunks, -1);
    std::vector<long long> prev_of(n_chunks, -1);  // previous chunk that used the same slot
    {
        long long last_in_slot[N_SLOTS];
        for (int s = 0; s < N_SLOTS; ++s) last_in_slot[s] = -1;
        int k = 0;
        for (size_t j = 0; j < n_chunks; ++j) {
            if (chunks[j].direct) continue;
            const int s = k++ % N_SLOTS;
            slot_of[j] = s;
            prev_of[j] = last_in_slot[s];
            last_in_slot[s] = (long long)j;
        }
    }
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    const int device = ctx->device;
    unsigned hw = std::thread::hardware_concurrency();
    int workers = hw ? (int)(hw / 2) : 4;
    if (workers > MAX_WORKERS) workers = MAX_WORKERS;
    if (workers < 1) workers = 1;
    if ((size_t)workers > n_chunks) workers = (int)n_chunks;
    auto work = [&]() {
        cudaSetDevice(device);
        while (true) {
            const size_t j = next.fetch_add(1);
            if (j >= n_chunks) return;
            if (chunks[j].direct) {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    staged[j] = 1;
                }
                cv.notify_all();
                continue;
            }
            const int s = slot_of[j];
            bool wait_drain = R->used[s];
            if (prev_of[j] >= 0) {  // the slot's previous chunk must have been issued ...
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return issued[prev_of[j]] || failed.load(); });
                wait_drain = true;
            }
            if (failed.load()) return;
            if (wait_drain && cudaEventSynchronize(R->drained[s]) != cudaSuccess) {  // ... and have left the host
                failed.store(1);
                cv.notify_all();
                return;
            }
            memcpy(R->slot[s], chunks[j].src, chunks[j].n);
            {
                std::lock_guard<std::mutex> lk(mu);
                staged[j] = 1;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(workers);
    for (int w = 0; w < workers; ++w) pool.emplace_back(work);
    cudaError_t err = cudaSuccess;
    for (size_t j = 0; j < n_chunks; ++j) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return staged[j] || failed.load(); });
        }
        if (failed.load()) break;
        const int s = slot_of[j];
        if (chunks[j].direct) {
            err = cudaMemcpyAsync(chunks[j].dst, chunks[j].src, chunks[j].n, cudaMemcpyHostToDevice, st);
        } else {
            err = cudaMemcpyAsync(chunks[j].dst, R->slot[s], chunks[j].n, cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaEventRecord(R->drained[s], st);
        }
        if (err != cudaSuccess) {
            failed.store(1);
            cv.notify_all();
            break;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            issued[j] = 1;
        }
        cv.notify_all();
        if (on_piece && last_chunk_of[j] >= 0) on_piece(last_chunk_of[j], user);  // piece fully enqueued on `stream`
    }
    for (auto& t : pool) t.join();
    for (size_t j = 0; j < n_chunks; ++j)
        if (slot_of[j] >= 0) R->used[slot_of[j]] = true;
    if (err != cudaSuccess) return m3d_fail(M3D_ERR_CUDA, "m3d_upload: %s", cudaGetErrorString(err));
    if (failed.load()) return m3d_fail(M3D_ERR_CUDA, "m3d_upload: staging failed");
    return M3D_OK;  // the tail of the copies is still in flight on `stream`
}

extern "C" int m3d_upload_batch(m3d_ctx* ctx, int n_pieces, const void* const* src_host, void* const* dst_dev,
                                const int64_t* n_bytes, void* stream) {
    return m3d_upload_batch_cb(ctx, n_pieces, src_host, dst_dev, n_bytes, stream, nullptr, nullptr);
}

extern "C" int m3d_upload(m3d_ctx* ctx, const void* src_host, void* dst_dev, int64_t n_bytes, void* stream) {
    return m3d_upload_batch(ctx, 1, &src_host, &dst_dev, &n_bytes, stream);
}
