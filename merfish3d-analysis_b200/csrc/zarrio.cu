// zarrio.cu -- the image store the decode stage reads, straight into HBM.
//
// The reference keeps every bit volume as an OME-NGFF v0.5 image: a Zarr v3 array of (16, 512, 512) chunks,
// each a Blosc-1 frame of zstd blocks over bit-shuffled uint16 / float32 (qi2labDataStore.py:1425-1529), and
// `_load_bit_data` (PD:1861-1874) has tensorstore decode them into NumPy arrays that are then copied to the
// GPU.  Here the host only does what must be sequential -- read the chunk file and entropy-decode its blocks,
// a pool of threads writing straight into page-locked slots -- and everything that is data-parallel happens on
// the device behind the DMA of the slot: undoing the bit / byte shuffle and placing the chunk (cropped at the
// volume edge and to the requested z window) into the tile's (z, y, x) volume.  No full-size host array, no
// host-side bit transposition, no pageable copy.
//
// The entropy coders are the system's libzstd.so.1 / liblz4.so.1, bound at run time (no headers needed for
// their stable C ABI).
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "zstd_decode.cuh"
#include "zstd_lanes.cuh"

namespace {

// ------------------------------------------------------------------------------------------ entropy coders
struct HostCodecs {
    size_t (*zstd_decompress)(void*, size_t, const void*, size_t) = nullptr;
    size_t (*zstd_compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*zstd_bound)(size_t) = nullptr;
    unsigned (*zstd_is_error)(size_t) = nullptr;
    void* (*zstd_create_dctx)() = nullptr;
    size_t (*zstd_free_dctx)(void*) = nullptr;
    size_t (*zstd_decompress_dctx)(void*, void*, size_t, const void*, size_t) = nullptr;
    int (*lz4_decompress)(const char*, char*, int, int) = nullptr;
    int (*lz4_compress)(const char*, char*, int, int) = nullptr;
    int (*lz4_bound)(int) = nullptr;
};

const HostCodecs& codecs() {
    static HostCodecs c;
    static std::once_flag once;
    std::call_once(once, [] {
        if (void* z = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL)) {
            c.zstd_decompress = reinterpret_cast<decltype(c.zstd_decompress)>(dlsym(z, "ZSTD_decompress"));
            c.zstd_compress = reinterpret_cast<decltype(c.zstd_compress)>(dlsym(z, "ZSTD_compress"));
            c.zstd_bound = reinterpret_cast<decltype(c.zstd_bound)>(dlsym(z, "ZSTD_compressBound"));
            c.zstd_is_error = reinterpret_cast<decltype(c.zstd_is_error)>(dlsym(z, "ZSTD_isError"));
            c.zstd_create_dctx = reinterpret_cast<decltype(c.zstd_create_dctx)>(dlsym(z, "ZSTD_createDCtx"));
            c.zstd_free_dctx = reinterpret_cast<decltype(c.zstd_free_dctx)>(dlsym(z, "ZSTD_freeDCtx"));
            c.zstd_decompress_dctx = reinterpret_cast<decltype(c.zstd_decompress_dctx)>(dlsym(z, "ZSTD_decompressDCtx"));
        }
        if (void* l = dlopen("liblz4.so.1", RTLD_NOW | RTLD_LOCAL)) {
            c.lz4_decompress = reinterpret_cast<decltype(c.lz4_decompress)>(dlsym(l, "LZ4_decompress_safe"));
            c.lz4_compress = reinterpret_cast<decltype(c.lz4_compress)>(dlsym(l, "LZ4_compress_default"));
            c.lz4_bound = reinterpret_cast<decltype(c.lz4_bound)>(dlsym(l, "LZ4_compressBound"));
        }
    });
    return c;
}

// One zstd frame -> dst.  A chunk is 32 frames of 256 KiB: the decompression context (~100 KB of tables) is kept
// per thread instead of being allocated and freed by every ZSTD_decompress call.
struct ThreadDCtx {
    void* ctx = nullptr;
    ~ThreadDCtx() {
        if (ctx) codecs().zstd_free_dctx(ctx);
    }
};

size_t zstd_decode_frame(const HostCodecs& C, void* dst, size_t cap, const void* src, size_t n) {
    if (C.zstd_create_dctx && C.zstd_free_dctx && C.zstd_decompress_dctx) {
        static thread_local ThreadDCtx t;
        if (!t.ctx) t.ctx = C.zstd_create_dctx();
        if (t.ctx) return C.zstd_decompress_dctx(t.ctx, dst, cap, src, n);
    }
    return C.zstd_decompress(dst, cap, src, n);
}

// ------------------------------------------------------------------------------------------ Blosc-1 frames
constexpr int BLOSC_HEADER = 16;
constexpr int FLAG_SHUFFLE = 0x1, FLAG_MEMCPY = 0x2, FLAG_BITSHUFFLE = 0x4, FLAG_DONT_SPLIT = 0x10;
constexpr int BLOSC_LZ4 = 1, BLOSC_ZSTD = 4;
constexpr int MAX_SPLITS = 16, MIN_BUFFERSIZE = 128;

struct BloscHeader {
    int version, versionlz, flags, typesize, codec;
    int64_t nbytes, blocksize, cbytes;
};

inline uint32_t le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline void put32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}

bool parse_blosc_header(const uint8_t* p, size_t n, BloscHeader& h) {
    if (n < (size_t)BLOSC_HEADER) return false;
    h.version = p[0];
    h.versionlz = p[1];
    h.flags = p[2];
    h.typesize = p[3];
    h.nbytes = le32(p + 4);
    h.blocksize = le32(p + 8);
    h.cbytes = le32(p + 12);
    h.codec = (h.flags >> 5) & 7;
    return h.typesize >= 1 && (h.nbytes == 0 || h.blocksize >= 1) && (size_t)h.cbytes <= n;
}

// ------------------------------------------------------------------------------------------ LZ4 blocks
// One LZ4 block (the raw block format Blosc stores: token, literal run, 2-byte match offset, match run) decoded
// by a team of lanes.  The parse is sequential and identical on every lane (all lanes read the same bytes);
// the two copies are what the team shares.  A match may overlap its own output (offset < length): byte i of it
// is byte (i mod offset) of the `offset` bytes before the match, which are all written before it starts, so the
// lanes never wait for each other inside a copy.  HostLanes (one lane) runs the same parse on the CPU -- it is the
// library's only LZ4 decoder, so the CPU tests exercise the code the GPU runs.  Every length is checked against
// both buffers: the input is a file from disk.
struct HostLanes {
    static inline void copy(uint8_t* d, const uint8_t* s, int64_t n) { memcpy(d, s, (size_t)n); }
    static inline void match(uint8_t* d, int64_t off, int64_t n) {
        if (off >= n) memcpy(d, d - off, (size_t)n);
        else for (int64_t i = 0; i < n; ++i) d[i] = d[i - off];
    }
    static inline void sync() {}
};

struct WarpLanes {
    static __device__ inline void copy(uint8_t* d, const uint8_t* s, int64_t n) {
        for (int64_t i = threadIdx.x & 31; i < n; i += 32) d[i] = s[i];
    }
    static __device__ inline void match(uint8_t* d, int64_t off, int64_t n) {
        const uint8_t* src = d - off;
        if (off >= n) {
            for (int64_t i = threadIdx.x & 31; i < n; i += 32) d[i] = src[i];
        } else {
            const uint32_t o = (uint32_t)off;  // off < 65536
            for (uint32_t i = threadIdx.x & 31; i < (uint32_t)n; i += 32) d[i] = src[i % o];  // n < 2^31
        }
    }
    static __device__ inline void sync() { __syncwarp(); }
};

template <class L>
__host__ __device__ inline bool lz4_decode_block(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_len) {
    int64_t ip = 0, op = 0;
    while (true) {
        if (ip >= in_len) return false;  // a block ends with a literals-only sequence
        const unsigned token = in[ip++];
        int64_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= in_len) return false;
                b = in[ip++];
                lit += b;
            } while (b == 255);
        }
        if (lit > in_len - ip || lit > out_len - op) return false;
        L::copy(out + op, in + ip, lit);
        ip += lit;
        op += lit;
        if (ip == in_len) break;
        if (in_len - ip < 2) return false;
        const int64_t off = (int64_t)in[ip] | ((int64_t)in[ip + 1] << 8);
        ip += 2;
        if (off == 0 || off > op) return false;
        int64_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (ip >= in_len) return false;
                b = in[ip++];
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (ml > out_len - op) return false;
        L::sync();  // the literals above and the previous match are visible to every lane
        L::match(out + op, off, ml);
        op += ml;
    }
    return op == out_len;
}

__host__ __device__ inline uint32_t le32_hd(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Up to FRAME_BATCH_MAX chunks per launch: the frames of several slots decoded by ONE grid.  A chunk is only 64 warps, and
// the device runs at most 32 streams' kernels side by side (hardware work queues), so one launch per chunk leaves most of
// the GPU idle however many slots are in flight; a launch over G chunks puts G x 64 latency chains behind each queue.
constexpr int FRAME_BATCH_MAX = 8;
struct FrameBatch {
    const uint8_t* frame[FRAME_BATCH_MAX];
    int64_t frame_len[FRAME_BATCH_MAX];
    uint8_t* out[FRAME_BATCH_MAX];
    uint8_t* lit[FRAME_BATCH_MAX];  // zstd only: literal scratch of the chunk
    int splits[FRAME_BATCH_MAX];
    int first_block[FRAME_BATCH_MAX + 1];  // thread blocks of chunk g: [first_block[g], first_block[g + 1])
    int n;
};

// A Blosc-1 frame of LZ4 streams, decoded on the device: one warp per stream (block j, split s).  `frame` is the
// chunk file as it sits on disk (the host only checked its 16-byte header); the warp finds its stream by
// walking the block's length prefixes, then decodes -- or copies a stored stream -- into its place of the
// still-shuffled chunk image `out`.  Anything inconsistent sets *error and the warp leaves.
__global__ void __launch_bounds__(128) blosc_lz4_decode_kernel(const __grid_constant__ FrameBatch batch,
                                                                int* __restrict__ error) {
    int g = 0;
    while (g + 1 < batch.n && (int)blockIdx.x >= batch.first_block[g + 1]) ++g;
    const uint8_t* frame = batch.frame[g];
    const int64_t frame_len = batch.frame_len[g];
    uint8_t* out = batch.out[g];
    const int splits_per_block = batch.splits[g];
    const int64_t warp = ((int64_t)((int)blockIdx.x - batch.first_block[g]) * blockDim.x + threadIdx.x) >> 5;
    const int typesize = frame[3], flags = frame[2];
    const int64_t nbytes = le32_hd(frame + 4), blocksize = le32_hd(frame + 8);
    const int64_t nblocks = (nbytes + blocksize - 1) / blocksize, leftover = nbytes % blocksize;
    const int64_t j = warp / splits_per_block;
    const int s = (int)(warp - j * splits_per_block);
    if (j >= nblocks) return;
    const bool is_left = (j == nblocks - 1) && leftover > 0;
    const int64_t bsize = is_left ? leftover : blocksize;
    const int nsplits = (!(flags & FLAG_DONT_SPLIT) && typesize <= MAX_SPLITS && bsize / typesize >= MIN_BUFFERSIZE &&
                         !is_left) ? typesize : 1;
    if (s >= nsplits) return;
    const int64_t neblock = bsize / nsplits;
    bool ok = BLOSC_HEADER + 4 * nblocks <= frame_len;
    int64_t pos = ok ? le32_hd(frame + BLOSC_HEADER + 4 * j) : 0;
    int64_t cb = 0;
    for (int t = 0; ok && t <= s; ++t) {  // length prefixes of the streams before ours, then ours
        ok = pos >= BLOSC_HEADER && pos + 4 <= frame_len;
        if (!ok) break;
        cb = (int32_t)le32_hd(frame + pos);
        pos += 4;
        ok = cb >= 0 && pos + cb <= frame_len;
        if (ok && t < s) pos += cb;
    }
    uint8_t* dst = out + j * blocksize + (int64_t)s * neblock;
    if (ok) {
        if (cb == neblock) WarpLanes::copy(dst, frame + pos, cb);
        else ok = lz4_decode_block<WarpLanes>(frame + pos, cb, dst, neblock);
    }
    if (!ok && (threadIdx.x & 31) == 0) atomicExch(error, 1);
}

// The same for Blosc-zstd frames, first device version (off unless M3D_ZARR_GPU_ZSTD=1): one warp per stream, the
// stream's zstd frame decoded by lane 0 with m3d_zstd::decode_frame (zstd_decode.cuh; tables in shared memory, the
// literal buffer in `lit_scratch`), stored streams copied by the whole warp.  Correct by construction -- it is the
// function the CPU tests pin to libzstd -- but lane-serial: sharing the Huffman streams and the copies among the
// lanes is the next step.
__global__ void __launch_bounds__(64) blosc_zstd_decode_kernel(const uint8_t* frame, int64_t frame_len, uint8_t* out,
                                                               int splits_per_block, uint8_t* lit_scratch,
                                                               int* __restrict__ error) {
    __shared__ m3d_zstd::Work work[2];
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int typesize = frame[3], flags = frame[2];
    const int64_t nbytes = le32_hd(frame + 4), blocksize = le32_hd(frame + 8);
    const int64_t nblocks = (nbytes + blocksize - 1) / blocksize, leftover = nbytes % blocksize;
    const int64_t j = warp / splits_per_block;
    const int s = (int)(warp - j * splits_per_block);
    if (j >= nblocks) return;
    const bool is_left = (j == nblocks - 1) && leftover > 0;
    const int64_t bsize = is_left ? leftover : blocksize;
    const int nsplits = (!(flags & FLAG_DONT_SPLIT) && typesize <= MAX_SPLITS && bsize / typesize >= MIN_BUFFERSIZE &&
                         !is_left) ? typesize : 1;
    if (s >= nsplits) return;
    const int64_t neblock = bsize / nsplits;
    bool ok = BLOSC_HEADER + 4 * nblocks <= frame_len;
    int64_t pos = ok ? le32_hd(frame + BLOSC_HEADER + 4 * j) : 0;
    int64_t cb = 0;
    for (int t = 0; ok && t <= s; ++t) {
        ok = pos >= BLOSC_HEADER && pos + 4 <= frame_len;
        if (!ok) break;
        cb = (int32_t)le32_hd(frame + pos);
        pos += 4;
        ok = cb >= 0 && pos + cb <= frame_len;
        if (ok && t < s) pos += cb;
    }
    uint8_t* dst = out + j * blocksize + (int64_t)s * neblock;
    if (ok) {
        if (cb == neblock) {
            WarpLanes::copy(dst, frame + pos, cb);
        } else {
            int good = 0;
            if ((threadIdx.x & 31) == 0) {
                m3d_zstd::Work& w = work[(threadIdx.x >> 5) & 1];
                w.lit = lit_scratch + warp * (int64_t)m3d_zstd::MAX_BLOCK;
                good = m3d_zstd::decode_frame(w, frame + pos, cb, dst, neblock) == neblock;
            }
            ok = __shfl_sync(0xffffffffu, good, 0) != 0;
        }
    }
    if (!ok && (threadIdx.x & 31) == 0) atomicExch(error, 1);
}

// Second device version (M3D_ZARR_GPU_ZSTD=2; zstd_lanes.cuh): the warp works as a team -- one lane per Huffman
// stream, shared copies.  Pinned on the host through the one-lane policy; NOT yet run on a device.
__global__ void __launch_bounds__(64) blosc_zstd_decode_kernel_v2(const __grid_constant__ FrameBatch batch,
                                                                  int* __restrict__ error) {
    __shared__ m3d_zstd::Work work[2];
    __shared__ m3d_zstd::LitPlan plans[2];
    int g = 0;
    while (g + 1 < batch.n && (int)blockIdx.x >= batch.first_block[g + 1]) ++g;
    const uint8_t* frame = batch.frame[g];
    const int64_t frame_len = batch.frame_len[g];
    uint8_t* out = batch.out[g];
    uint8_t* lit_scratch = batch.lit[g];
    const int splits_per_block = batch.splits[g];
    const int64_t warp = ((int64_t)((int)blockIdx.x - batch.first_block[g]) * blockDim.x + threadIdx.x) >> 5;
    const int typesize = frame[3], flags = frame[2];
    const int64_t nbytes = le32_hd(frame + 4), blocksize = le32_hd(frame + 8);
    const int64_t nblocks = (nbytes + blocksize - 1) / blocksize, leftover = nbytes % blocksize;
    const int64_t j = warp / splits_per_block;
    const int s = (int)(warp - j * splits_per_block);
    if (j >= nblocks) return;
    const bool is_left = (j == nblocks - 1) && leftover > 0;
    const int64_t bsize = is_left ? leftover : blocksize;
    const int nsplits = (!(flags & FLAG_DONT_SPLIT) && typesize <= MAX_SPLITS && bsize / typesize >= MIN_BUFFERSIZE &&
                         !is_left) ? typesize : 1;
    if (s >= nsplits) return;
    const int64_t neblock = bsize / nsplits;
    bool ok = BLOSC_HEADER + 4 * nblocks <= frame_len;
    int64_t pos = ok ? le32_hd(frame + BLOSC_HEADER + 4 * j) : 0;
    int64_t cb = 0;
    for (int t = 0; ok && t <= s; ++t) {
        ok = pos >= BLOSC_HEADER && pos + 4 <= frame_len;
        if (!ok) break;
        cb = (int32_t)le32_hd(frame + pos);
        pos += 4;
        ok = cb >= 0 && pos + cb <= frame_len;
        if (ok && t < s) pos += cb;
    }
    uint8_t* dst = out + j * blocksize + (int64_t)s * neblock;
    if (ok) {
        if (cb == neblock) {
            WarpLanes::copy(dst, frame + pos, cb);
        } else {
            const int wi = (threadIdx.x >> 5) & 1;
            if ((threadIdx.x & 31) == 0) work[wi].lit = lit_scratch + warp * (int64_t)m3d_zstd::MAX_BLOCK;
            __syncwarp();
            ok = m3d_zstd::decode_frame_lanes<m3d_zstd::Warp32>(work[wi], plans[wi], frame + pos, cb, dst, neblock) == neblock;
        }
    }
    if (!ok && (threadIdx.x & 31) == 0) atomicExch(error, 1);
}

// how the decoded bytes of a chunk are still arranged when they reach the device
enum ShuffleMode { SH_NONE = 0, SH_BYTE = 1, SH_BIT = 2 };

// Entropy-decode every block of a Blosc frame into `out` (h.nbytes bytes), leaving the per-block shuffle in
// place; *mode says which un-shuffle the consumer still owes.  Every offset is bounds-checked: the frame is a
// file from disk.  Returns nullptr on success, else a static description of what is wrong.
const char* blosc_decode_blocks(const uint8_t* frame, size_t n, const BloscHeader& h, uint8_t* out, int* mode) {
    *mode = SH_NONE;
    if (h.nbytes == 0) return nullptr;
    if (h.flags & FLAG_MEMCPY) {
        if ((size_t)(BLOSC_HEADER + h.nbytes) > n) return "blosc: stored frame is truncated";
        memcpy(out, frame + BLOSC_HEADER, (size_t)h.nbytes);
        return nullptr;
    }
    if ((h.flags & FLAG_SHUFFLE) && h.typesize > 1) *mode = SH_BYTE;
    else if (h.flags & FLAG_BITSHUFFLE) *mode = SH_BIT;
    const HostCodecs& C = codecs();
    if (h.codec == BLOSC_ZSTD && !C.zstd_decompress) return "blosc: libzstd.so.1 is not available";
    if (h.codec != BLOSC_ZSTD && h.codec != BLOSC_LZ4) return "blosc: unsupported inner codec (zstd and lz4 only)";
    const int64_t nblocks = (h.nbytes + h.blocksize - 1) / h.blocksize;
    const int64_t leftover = h.nbytes % h.blocksize;
    if ((size_t)(BLOSC_HEADER + 4 * nblocks) > n) return "blosc: block index is truncated";
    for (int64_t j = 0; j < nblocks; ++j) {
        const bool is_left = (j == nblocks - 1) && leftover > 0;
        const int64_t bsize = is_left ? leftover : h.blocksize;
        const int nsplits = (!(h.flags & FLAG_DONT_SPLIT) && h.typesize <= MAX_SPLITS &&
                             bsize / h.typesize >= MIN_BUFFERSIZE && !is_left) ? h.typesize : 1;
        const int64_t neblock = bsize / nsplits;
        size_t pos = le32(frame + BLOSC_HEADER + 4 * j);
        uint8_t* dst = out + j * h.blocksize;
        for (int s = 0; s < nsplits; ++s) {
            if (pos + 4 > n) return "blosc: block start outside the frame";
            const int64_t cb = (int32_t)le32(frame + pos);
            pos += 4;
            if (cb < 0 || pos + (size_t)cb > n) return "blosc: stream runs past the frame";
            if (cb == neblock) {
                memcpy(dst, frame + pos, (size_t)cb);
            } else if (h.codec == BLOSC_ZSTD) {
                const size_t r = zstd_decode_frame(C, dst, (size_t)neblock, frame + pos, (size_t)cb);
                if (C.zstd_is_error(r) || r != (size_t)neblock) return "blosc: zstd stream is corrupt";
            } else {
                if (!lz4_decode_block<HostLanes>(frame + pos, cb, dst, neblock)) return "blosc: lz4 stream is corrupt";
            }
            pos += (size_t)cb;
            dst += neblock;
        }
    }
    return nullptr;
}

// 8 x 8 bit-matrix transpose (byte i, bit j) <-> (byte j, bit i), both LSB first
__host__ __device__ inline uint64_t transpose8x8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;
    x ^= t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull;
    x ^= t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull;
    x ^= t ^ (t << 28);
    return x;
}

// The 8 elements (typesize TS) of group k of one block, whatever the shuffle.  `blk` = the block's decoded
// bytes, n_elem its element count, n8 = n_elem rounded down to a multiple of 8 (the bit-shuffled part).
template <int TS>
__host__ __device__ inline void fetch_group(const uint8_t* blk, int64_t n_elem, int mode, int64_t k, uint8_t out[8][TS]) {
    const int64_t n8 = n_elem & ~(int64_t)7;
    if (mode == SH_BIT && 8 * k + 8 <= n8) {
        const int64_t row = n8 >> 3;
#pragma unroll
        for (int b = 0; b < TS; ++b) {
            uint64_t x = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) x |= (uint64_t)blk[(int64_t)(b * 8 + i) * row + k] << (8 * i);
            x = transpose8x8(x);
#pragma unroll
            for (int j = 0; j < 8; ++j) out[j][b] = (uint8_t)(x >> (8 * j));
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t e = 8 * k + j;
#pragma unroll
        for (int b = 0; b < TS; ++b) {
            uint8_t v = 0;
            if (e < n_elem) v = (mode == SH_BYTE) ? blk[(int64_t)b * n_elem + e] : blk[e * TS + b];
            out[j][b] = v;
        }
    }
}

struct ChunkGeom {
    int64_t nbytes;       // decoded bytes of the chunk
    int64_t blocksize;    // Blosc block size (== nbytes for the other codecs)
    int mode;             // ShuffleMode
    int64_t cshape[3];    // chunk shape (z, y, x)
    int64_t origin[3];    // where chunk element (0,0,0) lands in the destination (may be negative / beyond)
    int64_t dshape[3];    // destination volume
};

template <int TS>
__host__ __device__ inline void place_group(const uint8_t v[8][TS], int64_t e0, int n_valid, const ChunkGeom& g,
                                            uint8_t* dst) {
    const int64_t plane = g.cshape[1] * g.cshape[2];
    const int64_t cz = e0 / plane;
    const int64_t rem = e0 - cz * plane;
    const int64_t cy = rem / g.cshape[2];
    const int64_t cx = rem - cy * g.cshape[2];
    if (n_valid == 8 && cx + 8 <= g.cshape[2]) {  // one row of the chunk
        const int64_t z = g.origin[0] + cz, y = g.origin[1] + cy, x = g.origin[2] + cx;
        if (z < 0 || z >= g.dshape[0] || y < 0 || y >= g.dshape[1]) return;
        uint8_t* p = dst + ((z * g.dshape[1] + y) * g.dshape[2] + x) * TS;
        if (x >= 0 && x + 8 <= g.dshape[2] && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (TS == 2 || TS == 4)) {
            if (TS == 2) {
                uint4 w;
                w.x = v[0][0] | (v[0][1] << 8) | (v[1][0] << 16) | ((uint32_t)v[1][1] << 24);
                w.y = v[2][0] | (v[2][1] << 8) | (v[3][0] << 16) | ((uint32_t)v[3][1] << 24);
                w.z = v[4][0] | (v[4][1] << 8) | (v[5][0] << 16) | ((uint32_t)v[5][1] << 24);
                w.w = v[6][0] | (v[6][1] << 8) | (v[7][0] << 16) | ((uint32_t)v[7][1] << 24);
                *reinterpret_cast<uint4*>(p) = w;
            } else {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    w[j] = v[j][0] | (v[j][1 % TS] << 8) | (v[j][2 % TS] << 16) | ((uint32_t)v[j][3 % TS] << 24);
                reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            return;
        }
    }
    for (int j = 0; j < n_valid; ++j) {
        const int64_t e = e0 + j;
        const int64_t ez = e / plane;
        const int64_t er = e - ez * plane;
        const int64_t ey = er / g.cshape[2];
        const int64_t ex = er - ey * g.cshape[2];
        const int64_t z = g.origin[0] + ez, y = g.origin[1] + ey, x = g.origin[2] + ex;
        if (z < 0 || z >= g.dshape[0] || y < 0 || y >= g.dshape[1] || x < 0 || x >= g.dshape[2]) continue;
        uint8_t* p = dst + ((z * g.dshape[1] + y) * g.dshape[2] + x) * TS;
#pragma unroll
        for (int b = 0; b < TS; ++b) p[b] = v[j][b];
    }
}

// group t of the chunk: block t / groups_per_block, group t % groups_per_block of that block
template <int TS>
__host__ __device__ inline void unshuffle_place_group(const uint8_t* staged, const ChunkGeom& g, int64_t t, uint8_t* dst) {
    const int64_t nb = g.blocksize / TS;  // elements per full block
    const int64_t gpb = (nb + 7) >> 3;
    const int64_t j = t / gpb, k = t - j * gpb;
    const int64_t first = j * nb;
    const int64_t total = g.nbytes / TS;
    if (first >= total) return;
    const int64_t n_elem = (total - first < nb) ? total - first : nb;
    if (8 * k >= n_elem) return;
    uint8_t v[8][TS];
    fetch_group<TS>(staged + j * g.blocksize, n_elem, g.mode, k, v);
    const int64_t left = n_elem - 8 * k;
    place_group<TS>(v, first + 8 * k, left < 8 ? (int)left : 8, g, dst);
}

template <int TS>
__global__ void __launch_bounds__(256) zarr_unshuffle_place_kernel(const uint8_t* __restrict__ staged, ChunkGeom g,
                                                                   int64_t n_groups, uint8_t* __restrict__ dst) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_groups; t += (int64_t)gridDim.x * blockDim.x)
        unshuffle_place_group<TS>(staged, g, t, dst);
}

// a chunk that was never written holds the array's fill value
template <typename T>
__global__ void __launch_bounds__(256) zarr_fill_chunk_kernel(ChunkGeom g, T fill, T* __restrict__ dst) {
    const int64_t n = g.cshape[0] * g.cshape[1] * g.cshape[2];
    const int64_t plane = g.cshape[1] * g.cshape[2];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ez = e / plane, er = e - ez * plane, ey = er / g.cshape[2], ex = er - ey * g.cshape[2];
        const int64_t z = g.origin[0] + ez, y = g.origin[1] + ey, x = g.origin[2] + ex;
        if (z < 0 || z >= g.dshape[0] || y < 0 || y >= g.dshape[1] || x < 0 || x >= g.dshape[2]) continue;
        dst[(z * g.dshape[1] + y) * g.dshape[2] + x] = fill;
    }
}

inline int64_t group_count(const ChunkGeom& g, int ts) {
    const int64_t nb = g.blocksize / ts;
    const int64_t nblocks = (g.nbytes + g.blocksize - 1) / g.blocksize;
    return nblocks * ((nb + 7) >> 3);
}

// Host path (NumPy destinations).  Two passes per chunk: un-shuffle block by block into a linear copy of the chunk
// (bit shuffle: 64 elements at a time -- an 8 x 8 byte transpose gathers, for each of 8 groups, its byte from the 8
// bit rows, then the 8 x 8 bit transpose), then copy the chunk's x-runs into the destination, clipped.
inline void transpose_bytes8x8(uint64_t r[8]) {
    uint64_t t;
    for (int i = 0; i < 8; i += 2) {
        t = ((r[i] >> 8) ^ r[i + 1]) & 0x00FF00FF00FF00FFull;
        r[i + 1] ^= t;
        r[i] ^= t << 8;
    }
    for (int i = 0; i < 8; i += 4)
        for (int k = 0; k < 2; ++k) {
            t = ((r[i + k] >> 16) ^ r[i + k + 2]) & 0x0000FFFF0000FFFFull;
            r[i + k + 2] ^= t;
            r[i + k] ^= t << 16;
        }
    for (int i = 0; i < 4; ++i) {
        t = ((r[i] >> 32) ^ r[i + 4]) & 0x00000000FFFFFFFFull;
        r[i + 4] ^= t;
        r[i] ^= t << 32;
    }
}

// bytes of a and b alternating (a0 b0 a1 b1 ...): w[0] from their low halves, w[1] from their high halves
inline uint64_t spread_bytes(uint64_t x) {
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    return (x | (x << 8)) & 0x00FF00FF00FF00FFull;
}
inline void interleave_bytes(uint64_t a, uint64_t b, uint64_t w[2]) {
    w[0] = spread_bytes(a & 0xFFFFFFFFull) | (spread_bytes(b & 0xFFFFFFFFull) << 8);
    w[1] = spread_bytes(a >> 32) | (spread_bytes(b >> 32) << 8);
}

template <int TS>
void unshuffle_block_host(const uint8_t* blk, int64_t n_elem, int mode, uint8_t* lin) {
    if (mode == SH_NONE) {
        memcpy(lin, blk, (size_t)(n_elem * TS));
        return;
    }
    if (mode == SH_BYTE) {
        for (int b = 0; b < TS; ++b) {
            const uint8_t* plane = blk + (int64_t)b * n_elem;
            for (int64_t e = 0; e < n_elem; ++e) lin[e * TS + b] = plane[e];
        }
        return;
    }
    const int64_t n8 = n_elem & ~(int64_t)7, row = n8 >> 3;
    int64_t k = 0;
    for (; k + 8 <= row; k += 8) {  // groups k .. k+7 = elements 8k .. 8k+63
        uint64_t y[TS][8];
        for (int b = 0; b < TS; ++b) {
            uint64_t r[8];
            for (int i = 0; i < 8; ++i) memcpy(&r[i], blk + (int64_t)(b * 8 + i) * row + k, 8);
            transpose_bytes8x8(r);  // r[g]: byte i = bit row i of group k+g
            for (int g = 0; g < 8; ++g) y[b][g] = transpose8x8(r[g]);  // byte j = byte b of element 8(k+g)+j
        }
        uint8_t* o = lin + 8 * k * TS;
        for (int g = 0; g < 8; ++g) {  // element j of the group = bytes j of y[0..TS-1][g], interleaved
            if (TS == 1) {
                memcpy(o + g * 8, &y[0][g], 8);
            } else if (TS == 2) {
                uint64_t w[2];
                interleave_bytes(y[0][g], y[1 % TS][g], w);
                memcpy(o + g * 16, w, 16);
            } else if (TS == 4) {
                uint64_t a[2], c[2], w[4];
                interleave_bytes(y[0][g], y[2 % TS][g], a);  // (b0, b2) pairs
                interleave_bytes(y[1 % TS][g], y[3 % TS][g], c);  // (b1, b3) pairs
                interleave_bytes(a[0], c[0], w);
                interleave_bytes(a[1], c[1], w + 2);
                memcpy(o + g * 32, w, 32);
            } else {
                for (int j = 0; j < 8; ++j)
                    for (int b = 0; b < TS; ++b) o[(g * 8 + j) * TS + b] = (uint8_t)(y[b][g] >> (8 * j));
            }
        }
    }
    for (; 8 * k < n_elem; ++k) {  // the last groups of the rows and the elements past n8 (stored plain)
        uint8_t v[8][TS];
        fetch_group<TS>(blk, n_elem, mode, k, v);
        const int64_t left = n_elem - 8 * k;
        for (int j = 0; j < (left < 8 ? (int)left : 8); ++j)
            for (int b = 0; b < TS; ++b) lin[(8 * k + j) * TS + b] = v[j][b];
    }
}

template <int TS>
void unshuffle_place_host(const uint8_t* staged, const ChunkGeom& g, uint8_t* dst) {
    const int64_t total = g.nbytes / TS;
    const int64_t nb = g.blocksize / TS;
    const uint8_t* lin = staged;
    std::vector<uint8_t> tmp;
    if (g.mode != SH_NONE) {
        tmp.resize((size_t)g.nbytes);
        for (int64_t first = 0, j = 0; first < total; first += nb, ++j)
            unshuffle_block_host<TS>(staged + j * g.blocksize, total - first < nb ? total - first : nb, g.mode,
                                     tmp.data() + first * TS);
        lin = tmp.data();
    }
    // x-runs of the chunk that fall inside the destination
    const int64_t x0 = g.origin[2] < 0 ? -g.origin[2] : 0;
    int64_t x1 = g.cshape[2];
    if (g.origin[2] + x1 > g.dshape[2]) x1 = g.dshape[2] - g.origin[2];
    if (x1 <= x0) return;
    for (int64_t cz = 0; cz < g.cshape[0]; ++cz) {
        const int64_t z = g.origin[0] + cz;
        if (z < 0 || z >= g.dshape[0]) continue;
        for (int64_t cy = 0; cy < g.cshape[1]; ++cy) {
            const int64_t y = g.origin[1] + cy;
            if (y < 0 || y >= g.dshape[1]) continue;
            memcpy(dst + ((z * g.dshape[1] + y) * g.dshape[2] + g.origin[2] + x0) * TS,
                   lin + ((cz * g.cshape[1] + cy) * g.cshape[2] + x0) * TS, (size_t)((x1 - x0) * TS));
        }
    }
}

// ------------------------------------------------------------------------------------------ chunk sources
bool read_range(const char* path, int64_t offset, int64_t length, std::vector<uint8_t>& buf, bool* missing) {
    *missing = false;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        *missing = (errno == ENOENT);
        return *missing;
    }
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return false;
    }
    if (length < 0) {
        length = (int64_t)sb.st_size - offset;
        if (length < 0) length = 0;
    } else if (offset < 0 || offset + length > (int64_t)sb.st_size) {  // shard index entry past the end of the file
        close(fd);
        return false;
    }
    buf.resize((size_t)length);
    size_t got = 0;
    while (got < (size_t)length) {
        const ssize_t r = pread(fd, buf.data() + got, (size_t)length - got, (off_t)(offset + (int64_t)got));
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    return got == (size_t)length;
}

bool read_all(int fd, int64_t offset, int64_t length, std::vector<uint8_t>& buf) {
    buf.resize((size_t)length);
    int64_t got = 0;
    while (got < length) {
        const ssize_t r = pread(fd, buf.data() + got, (size_t)(length - got), (off_t)(offset + got));
        if (r <= 0) break;
        got += r;
    }
    return got == length;
}

struct Decoded {
    int mode = SH_NONE;
    int64_t blocksize = 0;
};

// entropy-decode one encoded chunk into `out` (expected bytes); the shuffle stays in place
struct Bytes {
    const uint8_t* p = nullptr;
    size_t n = 0;
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

const char* decode_chunk_bytes(const m3d_zarr_chunk& c, const Bytes& enc, int64_t expected, uint8_t* out, Decoded* d) {
    d->mode = SH_NONE;
    d->blocksize = expected;
    if (c.codec == M3D_ZARR_RAW) {
        if ((int64_t)enc.size() != expected) return "zarr: raw chunk has the wrong size";
        memcpy(out, enc.data(), (size_t)expected);
        return nullptr;
    }
    if (c.codec == M3D_ZARR_ZSTD) {
        const HostCodecs& C = codecs();
        if (!C.zstd_decompress) return "zarr: libzstd.so.1 is not available";
        const size_t r = zstd_decode_frame(C, out, (size_t)expected, enc.data(), enc.size());
        if (C.zstd_is_error(r) || r != (size_t)expected) return "zarr: zstd chunk is corrupt";
        return nullptr;
    }
    if (c.codec != M3D_ZARR_BLOSC) return "zarr: unknown chunk codec";
    BloscHeader h;
    if (!parse_blosc_header(enc.data(), enc.size(), h)) return "blosc: bad frame header";
    if (h.nbytes != expected) return "blosc: frame does not hold one whole chunk";
    int mode = SH_NONE;
    if (const char* err = blosc_decode_blocks(enc.data(), enc.size(), h, out, &mode)) return err;
    if (mode != SH_NONE) {
        if (h.typesize != c.elem_size) return "blosc: typesize differs from the array's element size";
        if (h.blocksize % h.typesize != 0) return "blosc: block size is not a whole number of elements";
    }
    d->mode = mode;
    // un-shuffled data is already linear: placing it block by block would drop the bytes of an element that
    // straddles a block edge when the block size is not a multiple of the array's element size
    d->blocksize = ((h.flags & FLAG_MEMCPY) || mode == SH_NONE) ? expected : h.blocksize;
    return nullptr;
}

int check_chunk(const m3d_zarr_chunk& c) {
    if (!c.path || !c.dst) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read: null path / destination");
    if (c.elem_size != 1 && c.elem_size != 2 && c.elem_size != 4 && c.elem_size != 8)
        return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read: element size %d", c.elem_size);
    for (int a = 0; a < 3; ++a)
        if (c.chunk_shape[a] < 1 || c.dst_shape[a] < 1) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read: empty shape");
    return M3D_OK;
}

ChunkGeom geom_of(const m3d_zarr_chunk& c, const Decoded& d) {
    ChunkGeom g;
    g.nbytes = c.chunk_shape[0] * c.chunk_shape[1] * c.chunk_shape[2] * c.elem_size;
    g.blocksize = d.blocksize > 0 ? d.blocksize : g.nbytes;
    g.mode = d.mode;
    for (int a = 0; a < 3; ++a) {
        g.cshape[a] = c.chunk_shape[a];
        g.origin[a] = c.origin[a];
        g.dshape[a] = c.dst_shape[a];
    }
    return g;
}

// ------------------------------------------------------------------------------------------ slot ring
struct ZarrRing {
    std::vector<void*> pinned;    // host slot: the decoded (still shuffled) chunk, or the chunk file itself (LZ4)
    std::vector<void*> dev;       // device slot: the still-shuffled chunk image the un-shuffle kernel reads
    std::vector<void*> dev_comp;  // device slot: the chunk file, for frames the GPU entropy-decodes
    std::vector<void*> dev_lit;   // device slot: zstd literal buffers (128 KiB per stream), allocated on first use
    std::vector<size_t> dev_lit_bytes;
    std::vector<cudaEvent_t> drained;
    std::vector<cudaStream_t> streams;  // one per slot: chunks in different slots overlap their copies and kernels
    std::vector<char> used;
    size_t slot_bytes = 0;
    cudaEvent_t begin = nullptr;
    int* d_error = nullptr;
    int* h_error = nullptr;
};

std::mutex g_zring_mu;
std::vector<std::pair<m3d_ctx*, ZarrRing*>> g_zrings;

ZarrRing* zring_of(m3d_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_zring_mu);
    for (auto& p : g_zrings)
        if (p.first == ctx) return p.second;
    ZarrRing* r = new ZarrRing();
    g_zrings.push_back({ctx, r});
    return r;
}

void zring_free_slots(ZarrRing* r) {
    for (size_t s = 0; s < r->pinned.size(); ++s) {
        if (r->streams[s]) cudaStreamSynchronize(r->streams[s]);
        if (r->pinned[s]) cudaFreeHost(r->pinned[s]);
        if (r->dev[s]) cudaFree(r->dev[s]);
        if (r->dev_comp[s]) cudaFree(r->dev_comp[s]);
        if (r->dev_lit[s]) cudaFree(r->dev_lit[s]);
        if (r->drained[s]) cudaEventDestroy(r->drained[s]);
        if (r->streams[s]) cudaStreamDestroy(r->streams[s]);
    }
    r->pinned.clear();
    r->dev.clear();
    r->dev_comp.clear();
    r->dev_lit.clear();
    r->dev_lit_bytes.clear();
    r->drained.clear();
    r->streams.clear();
    r->used.clear();
    r->slot_bytes = 0;
}

void zring_free(ZarrRing* r) {
    zring_free_slots(r);
    if (r->begin) cudaEventDestroy(r->begin);
    if (r->d_error) cudaFree(r->d_error);
    if (r->h_error) cudaFreeHost(r->h_error);
    r->begin = nullptr;
    r->d_error = r->h_error = nullptr;
}

int zring_ensure(ZarrRing* r, int n_slots, size_t slot_bytes) {
    if (!r->begin) {
        M3D_CUDA(cudaEventCreateWithFlags(&r->begin, cudaEventDisableTiming));
        M3D_CUDA(cudaMalloc(reinterpret_cast<void**>(&r->d_error), sizeof(int)));
        M3D_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&r->h_error), sizeof(int), cudaHostAllocDefault));
    }
    if (r->slot_bytes >= slot_bytes && (int)r->pinned.size() >= n_slots) return M3D_OK;
    if (slot_bytes < r->slot_bytes) slot_bytes = r->slot_bytes;
    if (n_slots < (int)r->pinned.size()) n_slots = (int)r->pinned.size();
    zring_free_slots(r);
    for (int s = 0; s < n_slots; ++s) {
        r->pinned.push_back(nullptr);
        r->dev.push_back(nullptr);
        r->dev_comp.push_back(nullptr);
        r->dev_lit.push_back(nullptr);
        r->dev_lit_bytes.push_back(0);
        r->drained.push_back(nullptr);
        r->streams.push_back(nullptr);
        r->used.push_back(0);
        M3D_CUDA(cudaHostAlloc(&r->pinned[s], slot_bytes, cudaHostAllocDefault));
        M3D_CUDA(cudaMalloc(&r->dev[s], slot_bytes));
        M3D_CUDA(cudaMalloc(&r->dev_comp[s], slot_bytes));
        M3D_CUDA(cudaEventCreateWithFlags(&r->drained[s], cudaEventDisableTiming));
        M3D_CUDA(cudaStreamCreateWithFlags(&r->streams[s], cudaStreamNonBlocking));
    }
    r->slot_bytes = slot_bytes;
    return M3D_OK;
}

int io_threads() {
    if (const char* e = getenv("M3D_IO_THREADS")) {
        const int v = atoi(e);
        if (v >= 1) return v > 64 ? 64 : v;
    }
    unsigned hw = std::thread::hardware_concurrency();
    int w = hw ? (int)hw : 8;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {  // torchrun: the node's cores are shared by its ranks
        const int ranks = atoi(e);
        if (ranks > 1) w = w / ranks > 4 ? w / ranks : 4;
    }
    return w > 32 ? 32 : w;
}

template <int TS>
void launch_unshuffle(m3d_ctx* ctx, const uint8_t* staged, const ChunkGeom& g, uint8_t* dst, cudaStream_t st) {
    const int64_t n = group_count(g, TS);
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    M3D_LAUNCH(ctx, KF_ZARR_UNSHUFFLE, st, zarr_unshuffle_place_kernel<TS><<<(unsigned)blocks, 256, 0, st>>>(staged, g, n, dst));
}

int launch_for(m3d_ctx* ctx, int ts, const uint8_t* staged, const ChunkGeom& g, void* dst, cudaStream_t st) {
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    switch (ts) {
        case 1: launch_unshuffle<1>(ctx, staged, g, d, st); break;
        case 2: launch_unshuffle<2>(ctx, staged, g, d, st); break;
        case 4: launch_unshuffle<4>(ctx, staged, g, d, st); break;
        default: launch_unshuffle<8>(ctx, staged, g, d, st); break;
    }
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

int launch_fill(m3d_ctx* ctx, const m3d_zarr_chunk& c, cudaStream_t st) {
    const ChunkGeom g = geom_of(c, Decoded());
    const int64_t n = g.cshape[0] * g.cshape[1] * g.cshape[2];
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    switch (c.elem_size) {
        case 1: M3D_LAUNCH(ctx, KF_ZARR_FILL, st, zarr_fill_chunk_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>(g, (uint8_t)c.fill_bits, reinterpret_cast<uint8_t*>(c.dst))); break;
        case 2: M3D_LAUNCH(ctx, KF_ZARR_FILL, st, zarr_fill_chunk_kernel<uint16_t><<<(unsigned)blocks, 256, 0, st>>>(g, (uint16_t)c.fill_bits, reinterpret_cast<uint16_t*>(c.dst))); break;
        case 4: M3D_LAUNCH(ctx, KF_ZARR_FILL, st, zarr_fill_chunk_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(g, (uint32_t)c.fill_bits, reinterpret_cast<uint32_t*>(c.dst))); break;
        default: M3D_LAUNCH(ctx, KF_ZARR_FILL, st, zarr_fill_chunk_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, st>>>(g, (unsigned long long)c.fill_bits, reinterpret_cast<unsigned long long*>(c.dst))); break;
    }
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

}  // namespace

void m3d_release_zarr_ring(m3d_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_zring_mu);
    for (size_t i = 0; i < g_zrings.size(); ++i) {
        if (g_zrings[i].first != ctx) continue;
        zring_free(g_zrings[i].second);
        delete g_zrings[i].second;
        g_zrings.erase(g_zrings.begin() + i);
        return;
    }
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" int m3d_blosc_info(const void* frame, int64_t n_bytes, int64_t out[6]) {
    BloscHeader h;
    if (!frame || !out || !parse_blosc_header(reinterpret_cast<const uint8_t*>(frame), (size_t)n_bytes, h))
        return m3d_fail(M3D_ERR_ARG, "m3d_blosc_info: not a Blosc-1 frame");
    out[0] = h.nbytes;
    out[1] = h.blocksize;
    out[2] = h.cbytes;
    out[3] = h.typesize;
    out[4] = h.flags;
    out[5] = h.codec;
    return M3D_OK;
}

extern "C" int m3d_blosc_decode_host(const void* frame, int64_t n_bytes, void* dst, int64_t dst_capacity) {
    BloscHeader h;
    const uint8_t* f = reinterpret_cast<const uint8_t*>(frame);
    if (!frame || !dst || !parse_blosc_header(f, (size_t)n_bytes, h))
        return m3d_fail(M3D_ERR_ARG, "m3d_blosc_decode_host: not a Blosc-1 frame");
    if (h.nbytes > dst_capacity) return m3d_fail(M3D_ERR_CAPACITY, "m3d_blosc_decode_host: destination too small");
    std::vector<uint8_t> tmp((size_t)h.nbytes);
    int mode = SH_NONE;
    if (const char* err = blosc_decode_blocks(f, (size_t)n_bytes, h, tmp.data(), &mode)) return m3d_fail(M3D_ERR_ARG, "%s", err);
    if (mode == SH_NONE) {
        memcpy(dst, tmp.data(), (size_t)h.nbytes);
        return M3D_OK;
    }
    // a flat run of elements: a 1 x 1 x n "chunk" placed at the origin of a 1 x 1 x n volume; trailing bytes
    // that do not make a whole element are never shuffled
    const int ts = h.typesize;
    if (ts != 1 && ts != 2 && ts != 4 && ts != 8) return m3d_fail(M3D_ERR_ARG, "m3d_blosc_decode_host: typesize %d", ts);
    if (h.blocksize % ts) return m3d_fail(M3D_ERR_ARG, "m3d_blosc_decode_host: block size is not whole elements");
    const int64_t n_el = h.nbytes / ts;
    ChunkGeom g;
    g.nbytes = n_el * ts;
    g.blocksize = h.blocksize;
    g.mode = mode;
    g.cshape[0] = g.cshape[1] = 1;
    g.cshape[2] = n_el;
    g.origin[0] = g.origin[1] = g.origin[2] = 0;
    g.dshape[0] = g.dshape[1] = 1;
    g.dshape[2] = n_el;
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    if (n_el > 0) switch (ts) {
            case 1: unshuffle_place_host<1>(tmp.data(), g, d); break;
            case 2: unshuffle_place_host<2>(tmp.data(), g, d); break;
            case 4: unshuffle_place_host<4>(tmp.data(), g, d); break;
            default: unshuffle_place_host<8>(tmp.data(), g, d); break;
        }
    memcpy(d + n_el * ts, tmp.data() + n_el * ts, (size_t)(h.nbytes - n_el * ts));
    return M3D_OK;
}

namespace {
// forward shuffles of one block (the writer; element planes / bit rows as blosc_decode_blocks expects them)
void shuffle_block(const uint8_t* src, int64_t bsize, int ts, int mode, uint8_t* out) {
    const int64_t n = bsize / ts;
    if (mode == SH_BYTE) {
        for (int64_t e = 0; e < n; ++e)
            for (int b = 0; b < ts; ++b) out[(int64_t)b * n + e] = src[e * ts + b];
        memcpy(out + n * ts, src + n * ts, (size_t)(bsize - n * ts));
        return;
    }
    const int64_t n8 = n & ~(int64_t)7, row = n8 >> 3;
    for (int64_t k = 0; k < row; ++k)
        for (int b = 0; b < ts; ++b) {
            uint64_t x = 0;
            for (int j = 0; j < 8; ++j) x |= (uint64_t)src[(8 * k + j) * ts + b] << (8 * j);
            x = transpose8x8(x);
            for (int i = 0; i < 8; ++i) out[(int64_t)(b * 8 + i) * row + k] = (uint8_t)(x >> (8 * i));
        }
    memcpy(out + n8 * ts, src + n8 * ts, (size_t)(bsize - n8 * ts));
}
}  // namespace

extern "C" int64_t m3d_blosc_encode_bound(int64_t n_bytes, int64_t blocksize) {
    if (n_bytes < 0) return -1;
    if (blocksize <= 0) blocksize = 256 * 1024;
    if (blocksize > n_bytes) blocksize = n_bytes;
    // the encoder rounds the block size down to whole elements (never below half of it); incompressible blocks
    // are stored, so a block costs at most its size + its index entry + its length word
    const int64_t half = blocksize / 2 > 0 ? blocksize / 2 : 1;
    return BLOSC_HEADER + n_bytes + (n_bytes / half + 2) * 8 + 64;
}

extern "C" int m3d_blosc_encode_host(const void* src, int64_t n_bytes, int typesize, int cname, int clevel, int shuffle,
                                     int64_t blocksize, void* dst, int64_t dst_capacity, int64_t* out_bytes) {
    if (!src || !dst || !out_bytes || n_bytes < 0 || n_bytes > 0x7fffffff - 64 || typesize < 1 || typesize > 255)
        return m3d_fail(M3D_ERR_ARG, "m3d_blosc_encode_host: bad argument");
    if (cname != BLOSC_ZSTD && cname != BLOSC_LZ4) return m3d_fail(M3D_ERR_ARG, "m3d_blosc_encode_host: zstd (4) or lz4 (1)");
    if (shuffle < SH_NONE || shuffle > SH_BIT) return m3d_fail(M3D_ERR_ARG, "m3d_blosc_encode_host: shuffle 0, 1 or 2");
    const HostCodecs& C = codecs();
    if (cname == BLOSC_ZSTD && !C.zstd_compress) return m3d_fail(M3D_ERR_STATE, "libzstd.so.1 is not available");
    if (cname == BLOSC_LZ4 && !C.lz4_compress) return m3d_fail(M3D_ERR_STATE, "liblz4.so.1 is not available");
    if (blocksize <= 0) {  // c-blosc's choice for the non-split high-ratio codecs at clevel 5: 256 KiB
        blocksize = 32 * 1024 * (cname == BLOSC_ZSTD ? 2 : 1) * 4;
    }
    if (blocksize > n_bytes) blocksize = n_bytes > 0 ? n_bytes : 1;
    if (blocksize > typesize) blocksize = blocksize / typesize * typesize;
    if (dst_capacity < BLOSC_HEADER + n_bytes + 8 * ((n_bytes + blocksize - 1) / blocksize) + 8)
        return m3d_fail(M3D_ERR_CAPACITY, "m3d_blosc_encode_host: destination too small");
    const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
    uint8_t* o = reinterpret_cast<uint8_t*>(dst);
    int flags = (cname << 5) | FLAG_DONT_SPLIT;
    if (shuffle == SH_BYTE) flags |= FLAG_SHUFFLE;
    if (shuffle == SH_BIT) flags |= FLAG_BITSHUFFLE;
    o[0] = 2;
    o[1] = 1;
    o[3] = (uint8_t)typesize;
    put32(o + 4, (uint32_t)n_bytes);
    put32(o + 8, (uint32_t)blocksize);
    if (n_bytes < MIN_BUFFERSIZE || clevel <= 0) {
        o[2] = (uint8_t)(flags | FLAG_MEMCPY);
        memcpy(o + BLOSC_HEADER, s, (size_t)n_bytes);
        put32(o + 12, (uint32_t)(n_bytes + BLOSC_HEADER));
        *out_bytes = n_bytes + BLOSC_HEADER;
        return M3D_OK;
    }
    o[2] = (uint8_t)flags;
    const int64_t nblocks = (n_bytes + blocksize - 1) / blocksize;
    int64_t pos = BLOSC_HEADER + 4 * nblocks;
    std::vector<uint8_t> tmp((size_t)blocksize), comp;
    const int level = clevel * 2 - 1 < 1 ? 1 : clevel * 2 - 1;
    for (int64_t j = 0; j < nblocks; ++j) {
        const int64_t bsize = (j == nblocks - 1) ? n_bytes - j * blocksize : blocksize;
        const uint8_t* blk = s + j * blocksize;
        const bool do_byte = shuffle == SH_BYTE && typesize > 1, do_bit = shuffle == SH_BIT && bsize >= typesize;
        if (do_byte || do_bit) {
            shuffle_block(blk, bsize, typesize, do_byte ? SH_BYTE : SH_BIT, tmp.data());
            blk = tmp.data();
        }
        int64_t cb;
        if (cname == BLOSC_ZSTD) {
            comp.resize(C.zstd_bound((size_t)bsize));
            const size_t r = C.zstd_compress(comp.data(), comp.size(), blk, (size_t)bsize, level);
            cb = C.zstd_is_error(r) ? bsize : (int64_t)r;
        } else {
            comp.resize((size_t)C.lz4_bound((int)bsize));
            const int r = C.lz4_compress(reinterpret_cast<const char*>(blk), reinterpret_cast<char*>(comp.data()), (int)bsize, (int)comp.size());
            cb = r <= 0 ? bsize : r;
        }
        put32(o + BLOSC_HEADER + 4 * j, (uint32_t)pos);
        if (cb >= bsize) {  // stored
            put32(o + pos, (uint32_t)bsize);
            memcpy(o + pos + 4, blk, (size_t)bsize);
            pos += 4 + bsize;
        } else {
            put32(o + pos, (uint32_t)cb);
            memcpy(o + pos + 4, comp.data(), (size_t)cb);
            pos += 4 + cb;
        }
    }
    put32(o + 12, (uint32_t)pos);
    *out_bytes = pos;
    return M3D_OK;
}

extern "C" int m3d_zstd_host(int compress, const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int level,
                             int64_t* out_bytes) {
    const HostCodecs& C = codecs();
    if (!C.zstd_compress || !C.zstd_decompress) return m3d_fail(M3D_ERR_STATE, "libzstd.so.1 is not available");
    if (!src || !dst || !out_bytes || n_bytes < 0) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_host: bad argument");
    if (compress && (size_t)dst_capacity < C.zstd_bound((size_t)n_bytes)) return m3d_fail(M3D_ERR_CAPACITY, "m3d_zstd_host: destination too small");
    const size_t r = compress ? C.zstd_compress(dst, (size_t)dst_capacity, src, (size_t)n_bytes, level)
                              : C.zstd_decompress(dst, (size_t)dst_capacity, src, (size_t)n_bytes);
    if (C.zstd_is_error(r)) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_host: corrupt frame or destination too small");
    *out_bytes = (int64_t)r;
    return M3D_OK;
}

// The library's own zstd frame decoder (zstd_decode.cuh), on the host: the test hook that pins it to libzstd before it
// is moved behind the PCIe link.  Not used by the product path yet.
extern "C" int m3d_zstd_decode_builtin(const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int64_t* out_bytes) {
    if (!src || !dst || !out_bytes || n_bytes < 0 || dst_capacity < 0) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_decode_builtin: bad argument");
    std::vector<uint8_t> lit((size_t)m3d_zstd::MAX_BLOCK);
    std::vector<uint8_t> ws(sizeof(m3d_zstd::Work));
    m3d_zstd::Work* w = reinterpret_cast<m3d_zstd::Work*>(ws.data());
    memset(w, 0, sizeof(*w));
    w->lit = lit.data();
    const int64_t r = m3d_zstd::decode_frame(*w, reinterpret_cast<const uint8_t*>(src), n_bytes, reinterpret_cast<uint8_t*>(dst), dst_capacity);
    if (r < 0) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_decode_builtin: corrupt or unsupported frame");
    *out_bytes = r;
    return M3D_OK;
}

// The team-of-lanes arrangement of the same decoder (zstd_lanes.cuh) with a team of one: what pins its logic on the host.
extern "C" int m3d_zstd_decode_builtin_lanes(const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int64_t* out_bytes) {
    if (!src || !dst || !out_bytes || n_bytes < 0 || dst_capacity < 0) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_decode_builtin_lanes: bad argument");
    std::vector<uint8_t> lit((size_t)m3d_zstd::MAX_BLOCK);
    std::vector<uint8_t> ws(sizeof(m3d_zstd::Work));
    m3d_zstd::Work* w = reinterpret_cast<m3d_zstd::Work*>(ws.data());
    memset(w, 0, sizeof(*w));
    w->lit = lit.data();
    m3d_zstd::LitPlan plan;
    memset(&plan, 0, sizeof(plan));
    const int64_t r = m3d_zstd::decode_frame_lanes<m3d_zstd::OneLane>(*w, plan, reinterpret_cast<const uint8_t*>(src), n_bytes,
                                                                     reinterpret_cast<uint8_t*>(dst), dst_capacity);
    if (r < 0) return m3d_fail(M3D_ERR_ARG, "m3d_zstd_decode_builtin_lanes: corrupt or unsupported frame");
    *out_bytes = r;
    return M3D_OK;
}

// Host destination: the same chunk decode, the un-shuffle and placement run on the calling threads.  This is the
// reference's `tensorstore.read().result()` (DS:2235-2267) for callers that want a NumPy array (metadata probes,
// the normalisation sampler); the decode path proper uses m3d_zarr_read_chunks.
extern "C" int m3d_zarr_read_chunks_host(int n_chunks, const m3d_zarr_chunk* chunks) {
    if (n_chunks < 0 || (n_chunks > 0 && !chunks)) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read_chunks_host: bad argument");
    for (int j = 0; j < n_chunks; ++j)
        if (int rc = check_chunk(chunks[j])) return rc;
    std::atomic<int> next{0};
    std::mutex mu;
    std::string first_error;
    int workers = io_threads();
    if (workers > n_chunks) workers = n_chunks;
    auto work = [&]() {
        std::vector<uint8_t> enc, dec;
        while (true) {
            const int j = next.fetch_add(1);
            if (j >= n_chunks) return;
            const m3d_zarr_chunk& c = chunks[j];
            bool missing = false;
            const char* err = nullptr;
            Decoded d;
            const int64_t expected = c.chunk_shape[0] * c.chunk_shape[1] * c.chunk_shape[2] * c.elem_size;
            if (c.codec == M3D_ZARR_ABSENT) missing = true;
            else if (!read_range(c.path, c.offset, c.length, enc, &missing)) err = "zarr: cannot read the chunk file";
            if (!err && missing) {  // fill value, element by element through the same placement
                dec.assign((size_t)expected, 0);
                for (int64_t e = 0; e < expected / c.elem_size; ++e) memcpy(dec.data() + e * c.elem_size, &c.fill_bits, (size_t)c.elem_size);
            } else if (!err) {
                dec.resize((size_t)expected);
                err = decode_chunk_bytes(c, Bytes{enc.data(), enc.size()}, expected, dec.data(), &d);
            }
            if (err) {
                std::lock_guard<std::mutex> lk(mu);
                if (first_error.empty()) first_error = std::string(err) + " (" + c.path + ")";
                continue;
            }
            const ChunkGeom g = geom_of(c, d);
            uint8_t* dst = reinterpret_cast<uint8_t*>(c.dst);
            switch (c.elem_size) {
                case 1: unshuffle_place_host<1>(dec.data(), g, dst); break;
                case 2: unshuffle_place_host<2>(dec.data(), g, dst); break;
                case 4: unshuffle_place_host<4>(dec.data(), g, dst); break;
                default: unshuffle_place_host<8>(dec.data(), g, dst); break;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < workers; ++w) pool.emplace_back(work);
    if (n_chunks > 0) work();
    for (auto& t : pool) t.join();
    if (!first_error.empty()) return m3d_fail(M3D_ERR_ARG, "%s", first_error.c_str());
    return M3D_OK;
}

extern "C" int m3d_zarr_read_chunks(m3d_ctx* ctx, int n_chunks, const m3d_zarr_chunk* chunks, void* stream,
                                    m3d_piece_callback on_piece, void* user) {
    if (!ctx || n_chunks < 0 || (n_chunks > 0 && !chunks)) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read_chunks: bad argument");
    if (n_chunks == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t slot_bytes = 0;
    for (int j = 0; j < n_chunks; ++j) {
        if (int rc = check_chunk(chunks[j])) return rc;
        if (j > 0 && chunks[j].piece < chunks[j - 1].piece) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read_chunks: pieces must be grouped in ascending order");
        const size_t b = (size_t)(chunks[j].chunk_shape[0] * chunks[j].chunk_shape[1] * chunks[j].chunk_shape[2]) * chunks[j].elem_size;
        if (b > slot_bytes) slot_bytes = b;
    }
    if (slot_bytes > ((size_t)1 << 30)) return m3d_fail(M3D_ERR_ARG, "m3d_zarr_read_chunks: chunk larger than 1 GiB");
    slot_bytes += slot_bytes / 64 + 4096;  // room for a Blosc frame that did not compress (header, index, prefixes)
    int workers = io_threads();
    // Slots in flight.  A host-decoded chunk holds its slot for the decode (~5 ms) plus a short drain; a chunk the
    // device decodes holds it for the file read plus its decode kernel (LZ4 ~1 ms, zstd ~4 ms with others running),
    // and it is the number of such kernels running side by side (one stream per slot) that fills the GPU.
    int n_slots = 3 * workers;
    bool batch_zstd = false;
    int batch_default = 1;
    {
        // What will decode the frames?  The first chunk's Blosc header tells.  When the device does, the host threads only
        // copy chunk files from the page cache into pinned slots, and more than ~10 of them slow each other down
        // (measured, 16 bits x 32 x 2048^2, store -> HBM decoded GB/s: lz4 76 with 16 threads / 48 slots, 97 with 10 / 32;
        // zstd 47 with 16 / 48, 53 with 6-12 / 48; profiles/r2_zstd_device.txt).
        const char* gz0 = getenv("M3D_ZARR_GPU_ZSTD");
        const bool dev_zstd = !(gz0 && *gz0 && atoi(gz0) == 0), dev_lz4 = getenv("M3D_ZARR_HOST_LZ4") == nullptr;
        int inner = -1;
        if (chunks[0].codec == M3D_ZARR_BLOSC) {
            const int fd0 = open(chunks[0].path, O_RDONLY);
            uint8_t head0[BLOSC_HEADER];
            BloscHeader h0;
            if (fd0 >= 0) {
                if (pread(fd0, head0, BLOSC_HEADER, (off_t)chunks[0].offset) == BLOSC_HEADER &&
                    parse_blosc_header(head0, (size_t)1 << 40, h0) && !(h0.flags & FLAG_MEMCPY))
                    inner = h0.codec;
                close(fd0);
            }
        }
        const bool on_device = (inner == BLOSC_LZ4 && dev_lz4) || (inner == BLOSC_ZSTD && dev_zstd);
        if (on_device && getenv("M3D_IO_THREADS") == nullptr && workers > 10) workers = 10;
        if (on_device) n_slots = inner == BLOSC_LZ4 ? 32 : 96;
        batch_zstd = on_device;  // (name kept: chunks per decode launch, both device codecs)
        batch_default = inner == BLOSC_LZ4 ? 2 : 4;
    }
    if (workers > n_chunks) workers = n_chunks;
    if (const char* e = getenv("M3D_ZARR_SLOTS")) {  // tuning hook: slots (= chunks in flight) independent of the threads
        const int v = atoi(e);
        if (v >= 1) n_slots = v > 128 ? 128 : v;
    }
    int batch = 1;  // chunks per decode launch (device-decoded frames)
    if (batch_zstd) {
        batch = batch_default;  // measured: zstd 4 / 96 slots, lz4 2 / 32 slots (profiles/r2_zstd_device.txt)
        if (const char* e = getenv("M3D_ZARR_BATCH")) {
            const int v = atoi(e);
            if (v >= 1) batch = v > FRAME_BATCH_MAX ? FRAME_BATCH_MAX : v;
        }
    }
    while (n_slots > 3 && (size_t)n_slots * slot_bytes > ((size_t)1 << 30)) --n_slots;
    if (n_slots < batch) batch = n_slots;
    n_slots -= n_slots % batch;  // a group's slots are consecutive
    ZarrRing* R = zring_of(ctx);
    if (int rc = zring_ensure(R, n_slots, (slot_bytes + 255) & ~(size_t)255)) return rc;
    if ((int)R->pinned.size() < n_slots) n_slots = (int)R->pinned.size();  // (a ring kept from a wider call has more)
    const size_t slot_cap = R->slot_bytes;

    // Chunk j uses slot j % n_slots (and that slot's stream) once chunk j - n_slots has been issued and the slot has
    // drained.  A worker leaves in the pinned slot either the entropy-decoded chunk (HOST: zstd, raw, stored) or the
    // chunk file itself (GPU_LZ4: the device decodes it, so compressed bytes are what crosses PCIe).
    enum Kind : char { HOST = 0, GPU_LZ4 = 1, MISSING = 2 };
    std::vector<char> staged((size_t)n_chunks, 0), issued((size_t)n_chunks, 0), kind((size_t)n_chunks, HOST);
    std::vector<Decoded> info((size_t)n_chunks);
    std::vector<int64_t> comp_len((size_t)n_chunks, 0);
    std::vector<int> splits((size_t)n_chunks, 1);
    std::vector<int64_t> frame_blocksize((size_t)n_chunks, 0);  // Blosc block size of frames the device decodes
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    std::string first_error;
    const int device = ctx->device;
    const bool gpu_lz4 = getenv("M3D_ZARR_HOST_LZ4") == nullptr;
    const char* gz = getenv("M3D_ZARR_GPU_ZSTD");
    // Blosc-zstd frames (the reference's default codec, DS:58-60): decoded on the device by a team of lanes per Blosc block
    // (mode 2, the default: 46.9 GB/s decoded per GPU against 27 GB/s for 16 host threads, profiles/r2_zstd_device.txt);
    // M3D_ZARR_GPU_ZSTD=0 keeps the entropy stage on host threads (libzstd), 1 selects the lane-serial first version
    const int gpu_zstd_mode = (gz && *gz) ? atoi(gz) : 2;
    if (gpu_zstd_mode < 0 || gpu_zstd_mode > 2 || (gz && *gz && gpu_zstd_mode == 0 && strcmp(gz, "0") != 0))
        return m3d_fail(M3D_ERR_ARG, "M3D_ZARR_GPU_ZSTD must be 0, 1 or 2 (got '%s')", gz);
    const bool gpu_zstd = gpu_zstd_mode != 0;
    std::vector<char> on_gpu_zstd((size_t)n_chunks, 0);
    const char* mm = getenv("M3D_ZARR_MMAP");
    const bool use_mmap = mm ? atoi(mm) != 0 : true;  // +10 % on the zstd path: one copy of the file less
    auto fail = [&](const std::string& what) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (first_error.empty()) first_error = what;
        }
        failed.store(1);
        cv.notify_all();
    };
    auto work = [&]() {
        cudaSetDevice(device);
        std::vector<uint8_t> enc;
        while (true) {
            const int j = next.fetch_add(1);
            if (j >= n_chunks || failed.load()) return;
            const m3d_zarr_chunk& c = chunks[j];
            const int s = j % n_slots;
            const int64_t expected = c.chunk_shape[0] * c.chunk_shape[1] * c.chunk_shape[2] * c.elem_size;
            char k = HOST;
            int fd = -1;
            int64_t length = c.length;
            if (c.codec == M3D_ZARR_ABSENT) {
                k = MISSING;
            } else {
                fd = open(c.path, O_RDONLY);
                if (fd < 0) {
                    if (errno != ENOENT) return fail(std::string("zarr: cannot open ") + c.path);
                    k = MISSING;
                } else {
                    struct stat sb;
                    if (fstat(fd, &sb) != 0) {
                        close(fd);
                        return fail(std::string("zarr: cannot stat ") + c.path);
                    }
                    if (length < 0) {
                        length = (int64_t)sb.st_size - c.offset;
                        if (length < 0) length = 0;
                    } else if (c.offset < 0 || c.offset + length > (int64_t)sb.st_size) {
                        // a shard index entry that points past the end of a truncated shard: reading the mapping
                        // would raise SIGBUS, so it is refused here like the pread paths refuse it
                        close(fd);
                        return fail(std::string("zarr: chunk byte range lies outside ") + c.path);
                    }
                }
            }
            auto read_into = [&](uint8_t* dst, int64_t from, int64_t n) {
                int64_t got = 0;
                while (got < n) {
                    const ssize_t r = pread(fd, dst + got, (size_t)(n - got), (off_t)(c.offset + from + got));
                    if (r <= 0) break;
                    got += r;
                }
                return got == n;
            };
            uint8_t head[BLOSC_HEADER];
            BloscHeader h;
            if (k != MISSING && (gpu_lz4 || gpu_zstd) && c.codec == M3D_ZARR_BLOSC && length >= BLOSC_HEADER &&
                (size_t)length + 8 <= slot_cap /* the device bit readers load whole words */ && read_into(head, 0, BLOSC_HEADER) &&
                parse_blosc_header(head, (size_t)length, h) &&
                ((gpu_lz4 && h.codec == BLOSC_LZ4) || (gpu_zstd && h.codec == BLOSC_ZSTD)) &&
                !(h.flags & FLAG_MEMCPY) && h.nbytes == expected && h.blocksize % h.typesize == 0 &&
                (h.typesize == c.elem_size || !(h.flags & (FLAG_SHUFFLE | FLAG_BITSHUFFLE))) &&
                BLOSC_HEADER + 4 * ((h.nbytes + h.blocksize - 1) / h.blocksize) <= length &&
                // the device zstd decoder needs 128 KiB of literal scratch per stream: frames cut into so many
                // blocks that this passes 256 MiB per slot are decoded by the host threads instead
                (h.codec != BLOSC_ZSTD ||
                 ((h.nbytes + h.blocksize - 1) / h.blocksize) * (int64_t)(h.typesize <= MAX_SPLITS ? h.typesize : 1) *
                         (int64_t)m3d_zstd::MAX_BLOCK <= ((int64_t)256 << 20)))
                k = GPU_LZ4;
            // the encoded chunk: mapped straight from the page cache (no copy; M3D_ZARR_MMAP=0 reads it instead)
            Bytes src;
            void* map = MAP_FAILED;
            size_t map_len = 0;
            if (k == HOST && use_mmap && length > 0) {
                const int64_t page = sysconf(_SC_PAGESIZE);
                const int64_t lo = c.offset / page * page;
                map_len = (size_t)(c.offset - lo + length);
                map = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, (off_t)lo);
                if (map != MAP_FAILED) src = Bytes{reinterpret_cast<const uint8_t*>(map) + (c.offset - lo), (size_t)length};
            }
            if (k == HOST && map == MAP_FAILED) {
                if (!read_all(fd, c.offset, length, enc)) {
                    close(fd);
                    return fail(std::string("zarr: cannot read ") + c.path);
                }
                src = Bytes{enc.data(), enc.size()};
            }
            struct Unmap {
                void* m;
                size_t n;
                ~Unmap() {
                    if (m != MAP_FAILED) munmap(m, n);
                }
            } unmap{map, map_len};
            if (k != MISSING) {  // the slot: its previous chunk must have been issued, and must have left it
                bool wait_drain = R->used[s];
                if (j >= n_slots) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return issued[j - n_slots] || failed.load(); });
                    wait_drain = true;
                }
                if (!failed.load() && wait_drain && cudaEventSynchronize(R->drained[s]) != cudaSuccess) {
                    close(fd);
                    return fail("zarr: slot event failed");
                }
                if (failed.load()) {
                    close(fd);
                    return;
                }
                uint8_t* slot = reinterpret_cast<uint8_t*>(R->pinned[s]);
                if (k == GPU_LZ4) {
                    if (!read_into(slot, 0, length)) {
                        close(fd);
                        return fail(std::string("zarr: cannot read ") + c.path);
                    }
                    comp_len[j] = length;
                    on_gpu_zstd[j] = h.codec == BLOSC_ZSTD;
                    splits[j] = ((h.flags & FLAG_DONT_SPLIT) || h.typesize > MAX_SPLITS) ? 1 : h.typesize;
                    info[j].mode = ((h.flags & FLAG_SHUFFLE) && h.typesize > 1) ? SH_BYTE
                                   : (h.flags & FLAG_BITSHUFFLE) ? SH_BIT : SH_NONE;
                    info[j].blocksize = info[j].mode == SH_NONE ? expected : h.blocksize;  // linear data: one block
                    frame_blocksize[j] = h.blocksize;
                } else if (const char* err = decode_chunk_bytes(c, src, expected, slot, &info[j])) {
                    close(fd);
                    return fail(std::string(err) + " (" + c.path + ")");
                }
            }
            if (fd >= 0) close(fd);
            {
                std::lock_guard<std::mutex> lk(mu);
                kind[j] = k;
                staged[j] = 1;
            }
            cv.notify_all();
        }
    };
    // the slot streams start behind whatever `stream` has done so far (the destination's allocation, its readers)
    M3D_CUDA(cudaMemsetAsync(R->d_error, 0, sizeof(int), st));
    M3D_CUDA(cudaEventRecord(R->begin, st));
    std::vector<char> joined((size_t)n_slots, 0);
    std::vector<std::thread> pool;
    pool.reserve(workers);
    for (int w = 0; w < workers; ++w) pool.emplace_back(work);
    int rc = M3D_OK;
    bool any_gpu = false;
    auto cuda_ok = [&](cudaError_t e) {
        if (e != cudaSuccess && rc == M3D_OK) rc = m3d_fail(M3D_ERR_CUDA, "m3d_zarr_read_chunks: %s", cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    // Chunks are issued in aligned groups of `batch` consecutive chunks (= consecutive slots) on the stream of the group's
    // first slot: the device-zstd frames of a group go through ONE decode launch (FrameBatch), everything else as before.
    for (int j0 = 0; j0 < n_chunks && rc == M3D_OK; j0 += batch) {
        const int j1 = j0 + batch < n_chunks ? j0 + batch : n_chunks;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] {
                if (failed.load()) return true;
                for (int j = j0; j < j1; ++j)
                    if (!staged[j]) return false;
                return true;
            });
        }
        if (failed.load()) break;
        const int s0 = j0 % n_slots;
        cudaStream_t q = R->streams[s0];
        if (!joined[s0]) {
            cuda_ok(cudaStreamWaitEvent(q, R->begin, 0));
            joined[s0] = 1;
        }
        // the group's compressed frames cross PCIe, then one launch per codec decodes them
        FrameBatch zb, lb;
        zb.n = lb.n = 0;
        zb.first_block[0] = lb.first_block[0] = 0;
        for (int j = j0; j < j1 && rc == M3D_OK; ++j) {
            if (kind[j] != GPU_LZ4 || (on_gpu_zstd[j] && gpu_zstd_mode != 2)) continue;
            const int s = j % n_slots;
            const ChunkGeom g = geom_of(chunks[j], info[j]);
            const int64_t warps = (g.nbytes + frame_blocksize[j] - 1) / frame_blocksize[j] * splits[j];
            if (on_gpu_zstd[j]) {
                const size_t need = (size_t)warps * m3d_zstd::MAX_BLOCK;
                if (R->dev_lit_bytes[s] < need) {
                    cuda_ok(cudaStreamSynchronize(q));
                    if (R->dev_lit[s]) cudaFree(R->dev_lit[s]);
                    R->dev_lit[s] = nullptr;
                    R->dev_lit_bytes[s] = 0;
                    if (cuda_ok(cudaMalloc(&R->dev_lit[s], need))) R->dev_lit_bytes[s] = need;
                }
            }
            if (rc == M3D_OK &&
                cuda_ok(cudaMemcpyAsync(R->dev_comp[s], R->pinned[s], (size_t)comp_len[j], cudaMemcpyHostToDevice, q))) {
                FrameBatch& fb = on_gpu_zstd[j] ? zb : lb;
                const int i = fb.n++;
                fb.frame[i] = reinterpret_cast<const uint8_t*>(R->dev_comp[s]);
                fb.frame_len[i] = comp_len[j];
                fb.out[i] = reinterpret_cast<uint8_t*>(R->dev[s]);
                fb.lit[i] = on_gpu_zstd[j] ? reinterpret_cast<uint8_t*>(R->dev_lit[s]) : nullptr;
                fb.splits[i] = splits[j];
                // thread blocks: two warps (zstd) / four warps (lz4) each
                fb.first_block[i + 1] = fb.first_block[i] + (int)(on_gpu_zstd[j] ? (warps + 1) / 2 : (warps + 3) / 4);
            }
        }
        if (rc == M3D_OK && zb.n > 0) {
            any_gpu = true;
            M3D_LAUNCH(ctx, KF_ZARR_ZSTD, q,
                       blosc_zstd_decode_kernel_v2<<<(unsigned)zb.first_block[zb.n], 64, 0, q>>>(zb, R->d_error));
            cuda_ok(cudaGetLastError());
        }
        if (rc == M3D_OK && lb.n > 0) {
            any_gpu = true;
            M3D_LAUNCH(ctx, KF_ZARR_LZ4, q,
                       blosc_lz4_decode_kernel<<<(unsigned)lb.first_block[lb.n], 128, 0, q>>>(lb, R->d_error));
            cuda_ok(cudaGetLastError());
        }
        for (int j = j0; j < j1 && rc == M3D_OK; ++j) {
            const m3d_zarr_chunk& c = chunks[j];
            const int s = j % n_slots;
            if (kind[j] == MISSING) {
                rc = launch_fill(ctx, c, q);
            } else {
                const ChunkGeom g = geom_of(c, info[j]);
                if (kind[j] == GPU_LZ4 && !(on_gpu_zstd[j] && gpu_zstd_mode != 2)) {
                    // decoded by the group's launch above
                } else if (kind[j] == GPU_LZ4) {  // zstd mode 1: the lane-serial first version, one launch per chunk
                    any_gpu = true;
                    const int64_t nblocks = (g.nbytes + frame_blocksize[j] - 1) / frame_blocksize[j];
                    const int64_t warps = nblocks * splits[j];
                    const size_t need = (size_t)warps * m3d_zstd::MAX_BLOCK;
                    if (R->dev_lit_bytes[s] < need) {
                        cuda_ok(cudaStreamSynchronize(q));
                        if (R->dev_lit[s]) cudaFree(R->dev_lit[s]);
                        R->dev_lit[s] = nullptr;
                        R->dev_lit_bytes[s] = 0;
                        if (cuda_ok(cudaMalloc(&R->dev_lit[s], need))) R->dev_lit_bytes[s] = need;
                    }
                    if (rc == M3D_OK &&
                        cuda_ok(cudaMemcpyAsync(R->dev_comp[s], R->pinned[s], (size_t)comp_len[j], cudaMemcpyHostToDevice, q))) {
                        M3D_LAUNCH(ctx, KF_ZARR_ZSTD, q,
                                   blosc_zstd_decode_kernel<<<(unsigned)((warps + 1) / 2), 64, 0, q>>>(
                                       reinterpret_cast<const uint8_t*>(R->dev_comp[s]), comp_len[j],
                                       reinterpret_cast<uint8_t*>(R->dev[s]), splits[j],
                                       reinterpret_cast<uint8_t*>(R->dev_lit[s]), R->d_error));
                        cuda_ok(cudaGetLastError());
                    }
                } else {
                    cuda_ok(cudaMemcpyAsync(R->dev[s], R->pinned[s], (size_t)g.nbytes, cudaMemcpyHostToDevice, q));
                }
                if (rc == M3D_OK) rc = launch_for(ctx, c.elem_size, reinterpret_cast<const uint8_t*>(R->dev[s]), g, c.dst, q);
            }
            if (rc == M3D_OK && cuda_ok(cudaEventRecord(R->drained[s], q))) {
                R->used[s] = 1;
                if (j == j1 - 1) cuda_ok(cudaStreamWaitEvent(st, R->drained[s], 0));  // `stream` is ordered behind every group
            }
            if (rc != M3D_OK) break;
            {
                std::lock_guard<std::mutex> lk(mu);
                issued[j] = 1;
            }
            cv.notify_all();
            if (on_piece && (j == n_chunks - 1 || chunks[j + 1].piece != c.piece)) on_piece(c.piece, user);
        }
        if (rc != M3D_OK) {
            failed.store(1);
            cv.notify_all();
            break;
        }
    }
    for (auto& t : pool) t.join();
    if (rc != M3D_OK) return rc;
    if (failed.load()) return m3d_fail(M3D_ERR_ARG, "%s", first_error.empty() ? "m3d_zarr_read_chunks failed" : first_error.c_str());
    if (any_gpu) {  // frames decoded on the device report corruption through a flag: collect it before returning
        M3D_CUDA(cudaMemcpyAsync(R->h_error, R->d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
        M3D_CUDA(cudaStreamSynchronize(st));
        if (*R->h_error) return m3d_fail(M3D_ERR_ARG, "blosc: compressed stream is corrupt (device decode)");
    }
    return M3D_OK;  // the tail of the copies / kernels is still in flight on `stream`
}
