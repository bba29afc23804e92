// zstd_decode.cuh -- a Zstandard frame decoder written to run on either side of the PCIe link (RFC 8878).
//
// Why: the reference's default chunk codec is Blosc { zstd } (qi2labDataStore.py:1465-1475).  LZ4 frames are already
// decoded on the GPU (zarrio.cu); for zstd the entropy stage still runs on host threads through the system libzstd,
// and that is what bounds the image store -> HBM path.  This file is the decoder
// that moves it: no allocation, no recursion, no library calls, all state in a caller-provided workspace, every
// read and write bounds-checked (the input is a file from disk).  Round 1 ships it on the HOST only -- it is checked
// there against frames produced by the system libzstd at several levels (tests/test_cpu_zarr_store.py) through
// m3d_zstd_decode_builtin -- and the product path keeps calling libzstd; the device kernel that calls the same
// functions (one warp per 256 KiB Blosc block, tables in shared memory) is the next step.
//
// Format summary (RFC 8878): frame = magic, frame header, blocks (raw | RLE | compressed), optional checksum.
// A compressed block = literals section (raw | RLE | Huffman-coded in 1 or 4 streams, tree new or repeated) +
// sequences section (count, three FSE tables each predefined | RLE | described | repeated, one backward bitstream of
// interleaved literal-length / offset / match-length codes with extra bits), executed against the output so far
// with the three-entry repeat-offset history.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define M3D_HD __host__ __device__
#else
#define M3D_HD
#endif

namespace m3d_zstd {

constexpr uint32_t MAGIC = 0xFD2FB528u;
constexpr int MAX_BLOCK = 128 * 1024;
// RFC 8878 4.2.1: "this specification limits its maximum Number_of_Bits to 11" (libzstd's encoder: LitHufLog 11).  The
// tables below are sized by the format's own maxima (accuracy log 9 for literal / match lengths, 8 for offsets, 6 for
// Huffman weights): the workspace of one block decoder is ~10 KB, which is what lets 20 of them share an SM's shared memory.
constexpr int HUF_MAX_BITS = 11;
constexpr int FSE_MAX_AL = 9;

template <int AL>
struct FseTableT {
    static constexpr int MAX_AL = AL;
    uint8_t sym[1 << AL];
    uint8_t nb[1 << AL];
    uint16_t base[1 << AL];
    int al;
    int valid;
};

struct Work {
    uint16_t huf_tab[1 << HUF_MAX_BITS];  // symbol | bits consumed << 8, indexed by the next huf_bits bits
    int huf_bits;
    int huf_valid;
    FseTableT<9> ll, ml;
    FseTableT<8> of;
    FseTableT<6> wt;
    uint32_t rep[3];
    int16_t freq[256];
    uint8_t weights[256];
    uint16_t rank_pos[HUF_MAX_BITS + 2];
    uint8_t* lit;  // MAX_BLOCK bytes, caller-provided
};

M3D_HD inline int highbit32(uint32_t v) {  // v > 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// k (<= 56) bits starting at bit `bitoff` of p[0..n), least significant first; bits past the end read as zero
M3D_HD inline uint64_t bits_le(const uint8_t* p, int64_t n, int64_t bitoff, int k) {
    if (k == 0) return 0;
    const int64_t byte = bitoff >> 3;
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) {
        const int64_t b = byte + i;
        if (b >= 0 && b < n) v |= (uint64_t)p[b] << (8 * i);
    }
    v >>= (bitoff & 7);
    return v & ((k >= 64) ? ~0ull : ((1ull << k) - 1));
}

// Backward bitstream (Huffman streams, FSE weight stream, sequences): the last byte carries a 1-bit end marker above
// the payload; fields are taken from the top down.  Reading below bit 0 yields zeros in the low positions, and pos
// goes negative -- the decoders use that to detect the end exactly like the format's reference decoder.
struct BackBits {
    const uint8_t* p;
    int64_t n;
    int64_t pos;  // bits still unread (may go negative)
    int ok;
};

M3D_HD inline BackBits back_open(const uint8_t* p, int64_t n) {
    BackBits b{p, n, 0, 0};
    if (n < 1 || p[n - 1] == 0) return b;
    b.pos = (n - 1) * 8 + highbit32(p[n - 1]);
    b.ok = 1;
    return b;
}

M3D_HD inline uint64_t back_read(BackBits& b, int k) {
    if (k == 0) return 0;
    b.pos -= k;
    if (b.pos >= 0) return bits_le(b.p, b.n, b.pos, k);
    const int64_t avail = k + b.pos;  // bits that really exist
    if (avail <= 0) return 0;
    return bits_le(b.p, b.n, 0, (int)avail) << (-b.pos);
}

// ---------------------------------------------------------------------------------------------- FSE
// Normalised counts of an FSE table description (forward bitstream).  Returns the bytes consumed, or -1.
M3D_HD inline int64_t fse_read_counts(const uint8_t* p, int64_t n, int max_al, int max_symbols, int16_t* freq, int* n_symbols,
                                      int* al_out) {
    if (n < 1) return -1;
    int64_t bit = 0;
    const int al = (int)bits_le(p, n, bit, 4) + 5;
    bit += 4;
    if (al > max_al) return -1;
    int remaining = 1 << al;
    int s = 0;
    while (remaining > 0 && s < max_symbols) {
        const int nbits = highbit32((uint32_t)(remaining + 1)) + 1;
        uint32_t val = (uint32_t)bits_le(p, n, bit, nbits);
        const uint32_t lower = (1u << (nbits - 1)) - 1;
        const uint32_t threshold = (1u << nbits) - 1 - (uint32_t)(remaining + 1);
        if ((val & lower) < threshold) {
            bit += nbits - 1;
            val &= lower;
        } else {
            bit += nbits;
            if (val > lower) val -= threshold;
        }
        const int proba = (int)val - 1;
        remaining -= proba < 0 ? -proba : proba;
        freq[s++] = (int16_t)proba;
        if (proba == 0) {
            int rep = (int)bits_le(p, n, bit, 2);
            bit += 2;
            while (true) {
                for (int i = 0; i < rep && s < max_symbols; ++i) freq[s++] = 0;
                if (rep != 3) break;
                rep = (int)bits_le(p, n, bit, 2);
                bit += 2;
            }
        }
        if ((bit + 7) / 8 > n) return -1;
    }
    if (remaining != 0 || s > max_symbols) return -1;
    *n_symbols = s;
    *al_out = al;
    return (bit + 7) / 8;
}

template <class Table>
M3D_HD inline bool fse_build(Table& t, const int16_t* freq, int n_symbols, int al) {
    if (al > Table::MAX_AL || n_symbols > 256) return false;
    const int size = 1 << al;
    uint16_t next[256];
    int high = size;
    for (int s = 0; s < n_symbols; ++s) {
        if (freq[s] == -1) {
            if (high <= 0) return false;
            t.sym[--high] = (uint8_t)s;
            next[s] = 1;
        }
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < n_symbols; ++s) {
        if (freq[s] <= 0) continue;
        next[s] = (uint16_t)freq[s];
        for (int i = 0; i < freq[s]; ++i) {
            t.sym[pos] = (uint8_t)s;
            do {
                pos = (pos + step) & mask;
            } while (pos >= high);
        }
    }
    if (pos != 0) return false;
    for (int i = 0; i < size; ++i) {
        const int s = t.sym[i];
        const uint32_t x = next[s]++;
        const int nb = al - highbit32(x);
        t.nb[i] = (uint8_t)nb;
        t.base[i] = (uint16_t)((x << nb) - (uint32_t)size);
    }
    t.al = al;
    t.valid = 1;
    return true;
}

template <class Table>
M3D_HD inline void fse_rle(Table& t, int symbol) {
    t.sym[0] = (uint8_t)symbol;
    t.nb[0] = 0;
    t.base[0] = 0;
    t.al = 0;
    t.valid = 1;
}

// ---------------------------------------------------------------------------------------------- Huffman
// Weights -> decoding table.  weights[0..n) are the explicit ones; the last symbol's weight completes the sum to a
// power of two.
M3D_HD inline bool huf_build(Work& w, int n) {
    uint32_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (w.weights[i] > HUF_MAX_BITS) return false;
        if (w.weights[i]) total += 1u << (w.weights[i] - 1);
    }
    if (total == 0 || n >= 256) return false;
    const int max_bits = highbit32(total) + 1;
    if (max_bits > HUF_MAX_BITS) return false;
    const uint32_t left = (1u << max_bits) - total;
    if (left & (left - 1)) return false;  // must be a power of two
    w.weights[n] = (uint8_t)(highbit32(left) + 1);
    const int n_sym = n + 1;
    // entries are laid out by increasing weight (longest codes first), symbols in order within a weight
    uint32_t count[HUF_MAX_BITS + 2] = {0};
    for (int i = 0; i < n_sym; ++i) count[w.weights[i]]++;
    uint32_t pos = 0;
    for (int wt = 1; wt <= max_bits; ++wt) {
        w.rank_pos[wt] = (uint16_t)pos;
        pos += count[wt] << (wt - 1);
    }
    if (pos != (1u << max_bits)) return false;
    for (int i = 0; i < n_sym; ++i) {
        const int wt = w.weights[i];
        if (!wt) continue;
        const uint32_t len = 1u << (wt - 1);
        const uint32_t at = w.rank_pos[wt];
        for (uint32_t k = 0; k < len; ++k) {
            w.huf_tab[at + k] = (uint16_t)((uint32_t)i | ((uint32_t)(max_bits + 1 - wt) << 8));
        }
        w.rank_pos[wt] = (uint16_t)(at + len);
    }
    w.huf_bits = max_bits;
    w.huf_valid = 1;
    return true;
}

// Huffman tree description at p; returns the bytes it occupies, or -1.
M3D_HD inline int64_t huf_read_tree(Work& w, const uint8_t* p, int64_t n) {
    if (n < 1) return -1;
    const int head = p[0];
    int n_weights = 0;
    int64_t used;
    if (head >= 128) {  // direct: 4 bits per weight
        n_weights = head - 127;
        used = 1 + (n_weights + 1) / 2;
        if (used > n) return -1;
        for (int i = 0; i < n_weights; ++i) {
            const uint8_t b = p[1 + i / 2];
            w.weights[i] = (i & 1) ? (b & 15) : (b >> 4);
        }
    } else {  // FSE-compressed weights, two interleaved states, read until the bitstream runs dry
        used = 1 + head;
        if (head < 1 || used > n) return -1;
        int n_sym = 0, al = 0;
        const int64_t hdr = fse_read_counts(p + 1, head, 6, 256, w.freq, &n_sym, &al);
        if (hdr < 0 || hdr >= head) return -1;
        if (!fse_build(w.wt, w.freq, n_sym, al)) return -1;
        BackBits b = back_open(p + 1 + hdr, head - hdr);
        if (!b.ok) return -1;
        uint32_t s1 = (uint32_t)back_read(b, al), s2 = (uint32_t)back_read(b, al);
        if (b.pos < 0) return -1;
        while (true) {
            if (n_weights >= 255) return -1;
            w.weights[n_weights++] = w.wt.sym[s1];
            s1 = w.wt.base[s1] + (uint32_t)back_read(b, w.wt.nb[s1]);
            if (b.pos < 0) {
                if (n_weights >= 255) return -1;
                w.weights[n_weights++] = w.wt.sym[s2];
                break;
            }
            if (n_weights >= 255) return -1;
            w.weights[n_weights++] = w.wt.sym[s2];
            s2 = w.wt.base[s2] + (uint32_t)back_read(b, w.wt.nb[s2]);
            if (b.pos < 0) {
                if (n_weights >= 255) return -1;
                w.weights[n_weights++] = w.wt.sym[s1];
                break;
            }
        }
    }
    if (!huf_build(w, n_weights)) return -1;
    return used;
}

// Window over a backward bitstream made of 32-bit words: word i covers stream bytes [4 i - off, 4 i - off + 4), bytes
// outside [0, n) read as zero.  On the device `off` aligns the words with the buffer (one aligned 32-bit load per 32
// consumed bits instead of eight byte loads per field); the host assembles the same words from bytes, so the logic the
// CPU tests pin is the logic the GPU runs.  The caller guarantees that the 8 bytes after the stream may be READ (they
// belong to the same frame / staging slot); their bits are never used.
struct BackWin {
    const uint8_t* p;
    int32_t n;
    int32_t off;   // 0..3: p - (p rounded down to 4 bytes); 0 on the host
    int32_t idx;   // lo = word(idx), hi = word(idx + 1)
    uint32_t lo, hi;
};

M3D_HD inline uint32_t backwin_word(const BackWin& b, int32_t i) {
    if (i < 0) return 0u;
#ifdef __CUDA_ARCH__
    uint32_t v = reinterpret_cast<const uint32_t*>(b.p - b.off)[i];
    if (i == 0) v &= 0xFFFFFFFFu << (8 * b.off);  // bytes before the stream
    return v;
#else
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) {
        const int64_t at = (int64_t)4 * i - b.off + k;
        if (at >= 0 && at < b.n) v |= (uint32_t)b.p[at] << (8 * k);
    }
    return v;
#endif
}

// the k (<= 25) bits [pos, pos + k) of the stream, pos possibly negative (bits below 0 read as zero)
M3D_HD inline uint32_t backwin_peek(BackWin& b, int32_t pos, int k) {
    const int32_t q = pos + 8 * b.off;
    const int32_t wi = q >> 5;  // floor
    if (wi != b.idx) {
        if (wi == b.idx - 1) {
            b.hi = b.lo;
            b.lo = backwin_word(b, wi);
        } else {
            b.lo = backwin_word(b, wi);
            b.hi = backwin_word(b, wi + 1);
        }
        b.idx = wi;
    }
    const int sh = q & 31;
#ifdef __CUDA_ARCH__
    const uint32_t v = __funnelshift_r(b.lo, b.hi, sh);
#else
    const uint32_t v = (uint32_t)((((uint64_t)b.hi << 32) | b.lo) >> sh);
#endif
    return v & ((1u << k) - 1u);
}

// One Huffman stream (RFC 8878 4.2.2): the state is the next `huf_bits` bits; each symbol consumes huf_nb[state] of them.
// Same results as the field-by-field reader it replaces (pinned against libzstd by the CPU tests), with one aligned word
// load per 32 bits and four symbols per 32-bit store where the destination allows.
M3D_HD inline bool huf_decode_stream(const Work& w, const uint8_t* p, int64_t n, uint8_t* out, int64_t n_out) {
    if (n < 1 || n > (1 << 26) || p[n - 1] == 0) return false;
    const int mb = w.huf_bits;
    const uint32_t mask = (1u << mb) - 1;
    BackWin b;
    b.p = p;
    b.n = (int32_t)n;
#ifdef __CUDA_ARCH__
    b.off = (int32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
#else
    b.off = 0;
#endif
    b.idx = INT32_MIN / 2;
    b.lo = b.hi = 0;
    int32_t pos = (int32_t)(n - 1) * 8 + highbit32(p[n - 1]);  // bits still unread (may go negative)
    pos -= mb;
    uint32_t state = (pos >= 0) ? backwin_peek(b, pos, mb)
                                : (pos + mb > 0 ? (backwin_peek(b, 0, pos + mb) << (-pos)) : 0u);
    int64_t i = 0;
    // ---- main loop: four symbols per round while at least 4 x 11 payload bits remain.  Nothing in the body depends on
    // the data except through predicated moves, so the lanes of a warp that decode different streams stay in lockstep
    // (an earlier version with early exits and a refill branch ran the four streams of a block one after the other).
    // While pos >= 0 every field lies inside the stream: no zero-fill, no dry-run check.
    while (i < n_out && (reinterpret_cast<uintptr_t>(out + i) & 3u) && pos >= HUF_MAX_BITS) {
        const uint32_t e = w.huf_tab[state];
        const int nb = (int)(e >> 8);
        out[i++] = (uint8_t)e;
        pos -= nb;
        state = ((state << nb) & mask) | (nb ? backwin_peek(b, pos, nb) : 0u);
    }
    {
        const int32_t base = 8 * b.off;
        int32_t q0 = pos + base;
        int32_t idx = q0 >> 5;
        uint32_t lo = backwin_word(b, idx), hi = backwin_word(b, idx + 1);
        while (i + 4 <= n_out && pos >= 4 * HUF_MAX_BITS) {
            uint32_t packed = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = w.huf_tab[state];
                const int nb = (int)(e >> 8);
                packed |= (e & 0xFFu) << (8 * u);
                pos -= nb;
                const int32_t q = pos + base;
                const int32_t wi = q >> 5;
                if (wi != idx) {  // at most one word further down (nb <= 11): predicated, not a branch
                    hi = lo;
#ifdef __CUDA_ARCH__
                    lo = reinterpret_cast<const uint32_t*>(b.p - b.off)[wi];
#else
                    lo = backwin_word(b, wi);
#endif
                    idx = wi;
                }
#ifdef __CUDA_ARCH__
                const uint32_t v = __funnelshift_r(lo, hi, q & 31);
#else
                const uint32_t v = (uint32_t)((((uint64_t)hi << 32) | lo) >> (q & 31));
#endif
                state = ((state << nb) & mask) | (v & ((1u << nb) - 1u));
            }
#ifdef __CUDA_ARCH__
            *reinterpret_cast<uint32_t*>(out + i) = packed;
#else
            memcpy(out + i, &packed, 4);
#endif
            i += 4;
        }
        b.idx = INT32_MIN / 2;  // the tail reloads its window
    }
    // ---- tail (the last few symbols, and streams shorter than the main loop's margin): checked field by field; fields
    // that reach below bit 0 read zeros in the missing low positions
    int bad = 0;
    for (; i < n_out; ++i) {
        bad |= (pos <= -mb);  // ran dry before the last symbol
        const uint32_t e = w.huf_tab[state];
        const int nb = (int)(e >> 8);
        out[i] = (uint8_t)e;
        uint32_t fresh = 0;
        if (nb) {
            pos -= nb;
            if (pos >= 0) fresh = backwin_peek(b, pos, nb);
            else if (pos + nb > 0) fresh = backwin_peek(b, 0, pos + nb) << (-pos);
        }
        state = ((state << nb) & mask) | fresh;
    }
    if (bad) return false;
    return pos == -mb;  // every payload bit used, nothing more
}

// ---------------------------------------------------------------------------------------------- literals
// Decodes the literals section at p[0..n) into w.lit; returns the section's size in bytes (or -1), *n_lit = count.
M3D_HD inline int64_t literals_decode(Work& w, const uint8_t* p, int64_t n, int64_t* n_lit) {
    if (n < 1) return -1;
    const int type = p[0] & 3, fmt = (p[0] >> 2) & 3;
    if (type < 2) {  // raw / RLE
        int64_t hdr, regen;
        if ((fmt & 1) == 0) {
            hdr = 1;
            regen = p[0] >> 3;
        } else if (fmt == 1) {
            if (n < 2) return -1;
            hdr = 2;
            regen = (p[0] >> 4) | ((int64_t)p[1] << 4);
        } else {
            if (n < 3) return -1;
            hdr = 3;
            regen = (p[0] >> 4) | ((int64_t)p[1] << 4) | ((int64_t)p[2] << 12);
        }
        if (regen > MAX_BLOCK) return -1;
        if (type == 0) {
            if (hdr + regen > n) return -1;
            memcpy(w.lit, p + hdr, (size_t)regen);
            *n_lit = regen;
            return hdr + regen;
        }
        if (hdr + 1 > n) return -1;
        memset(w.lit, p[hdr], (size_t)regen);
        *n_lit = regen;
        return hdr + 1;
    }
    int64_t hdr, regen, comp;
    int streams;
    if (fmt == 0 || fmt == 1) {
        if (n < 3) return -1;
        const uint32_t v = p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16);
        hdr = 3;
        streams = fmt == 0 ? 1 : 4;
        regen = (v >> 4) & 0x3FF;
        comp = (v >> 14) & 0x3FF;
    } else if (fmt == 2) {
        if (n < 4) return -1;
        const uint32_t v = p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        hdr = 4;
        streams = 4;
        regen = (v >> 4) & 0x3FFF;
        comp = (v >> 18) & 0x3FFF;
    } else {
        if (n < 5) return -1;
        const uint64_t v = (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24) |
                           ((uint64_t)p[4] << 32);
        hdr = 5;
        streams = 4;
        regen = (int64_t)((v >> 4) & 0x3FFFF);
        comp = (int64_t)((v >> 22) & 0x3FFFF);
    }
    if (regen > MAX_BLOCK || hdr + comp > n) return -1;
    const uint8_t* q = p + hdr;
    int64_t left = comp;
    if (type == 2) {
        const int64_t tree = huf_read_tree(w, q, left);
        if (tree < 0) return -1;
        q += tree;
        left -= tree;
    } else if (!w.huf_valid) {
        return -1;  // "treeless" block without a previous tree
    }
    if (streams == 1) {
        if (!huf_decode_stream(w, q, left, w.lit, regen)) return -1;
    } else {
        if (left < 6) return -1;
        const int64_t s1 = q[0] | (q[1] << 8), s2 = q[2] | (q[3] << 8), s3 = q[4] | (q[5] << 8);
        const int64_t s4 = left - 6 - s1 - s2 - s3;
        if (s4 < 1) return -1;
        const int64_t each = (regen + 3) / 4;
        const int64_t last = regen - 3 * each;
        if (last < 0) return -1;
        const uint8_t* s = q + 6;
        if (!huf_decode_stream(w, s, s1, w.lit, each)) return -1;
        if (!huf_decode_stream(w, s + s1, s2, w.lit + each, each)) return -1;
        if (!huf_decode_stream(w, s + s1 + s2, s3, w.lit + 2 * each, each)) return -1;
        if (!huf_decode_stream(w, s + s1 + s2 + s3, s4, w.lit + 3 * each, last)) return -1;
    }
    *n_lit = regen;
    return hdr + comp;
}

// ---------------------------------------------------------------------------------------------- sequences
M3D_HD inline void predefined_counts(int which, int16_t* f, int* n, int* al) {
    // RFC 8878 section 3.1.1.3.2.2: default distributions of literal-length, match-length and offset codes
    if (which == 0) {
        const int8_t d[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
        for (int i = 0; i < 36; ++i) f[i] = d[i];
        *n = 36;
        *al = 6;
    } else if (which == 1) {
        const int8_t d[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
        for (int i = 0; i < 29; ++i) f[i] = d[i];
        *n = 29;
        *al = 5;
    } else {
        const int8_t d[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                              1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
        for (int i = 0; i < 53; ++i) f[i] = d[i];
        *n = 53;
        *al = 6;
    }
}

// one of the three code tables of a sequences section; which: 0 LL, 1 OF, 2 ML.  Returns bytes consumed or -1.
template <class Table>
M3D_HD inline int64_t seq_table(Work& w, Table& t, int which, int mode, const uint8_t* p, int64_t n) {
    const int max_al[3] = {9, 8, 9}, max_sym[3] = {36, 32, 53};
    if (mode == 0) {
        int ns, al;
        predefined_counts(which, w.freq, &ns, &al);
        return fse_build(t, w.freq, ns, al) ? 0 : -1;
    }
    if (mode == 1) {
        if (n < 1 || p[0] >= max_sym[which]) return -1;
        fse_rle(t, p[0]);
        return 1;
    }
    if (mode == 2) {
        int ns = 0, al = 0;
        const int64_t used = fse_read_counts(p, n, max_al[which], max_sym[which], w.freq, &ns, &al);
        if (used < 0 || !fse_build(t, w.freq, ns, al)) return -1;
        return used;
    }
    return t.valid ? 0 : -1;  // repeat
}

M3D_HD inline void ll_code(int c, uint32_t* base, int* nb) {
    const uint32_t b[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512,
                            1024, 2048, 4096, 8192, 16384, 32768, 65536};
    const uint8_t e[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    *base = b[c];
    *nb = e[c];
}

M3D_HD inline void ml_code(int c, uint32_t* base, int* nb) {
    const uint32_t b[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30,
                            31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
                            16387, 32771, 65539};
    const uint8_t e[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                           0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    *base = b[c];
    *nb = e[c];
}

// Sequences section at p[0..n): decode and execute against out[0..cap) whose first *op bytes are already written
// (earlier blocks of the frame are the match window).  Literals come from w.lit[0..n_lit).
M3D_HD inline bool sequences_execute(Work& w, const uint8_t* p, int64_t n, int64_t n_lit, uint8_t* out, int64_t cap, int64_t* op) {
    int64_t o = *op, lp = 0;
    if (n < 1) return false;
    int64_t nseq = p[0], at = 1;
    if (nseq >= 128) {
        if (nseq < 255) {
            if (n < 2) return false;
            nseq = ((nseq - 128) << 8) + p[1];
            at = 2;
        } else {
            if (n < 3) return false;
            nseq = p[1] + ((int64_t)p[2] << 8) + 0x7F00;
            at = 3;
        }
    }
    if (nseq > 0) {
        if (at >= n) return false;
        const int modes = p[at++];
        if (modes & 3) return false;
        int64_t used = seq_table(w, w.ll, 0, (modes >> 6) & 3, p + at, n - at);
        if (used < 0) return false;
        at += used;
        used = seq_table(w, w.of, 1, (modes >> 4) & 3, p + at, n - at);
        if (used < 0) return false;
        at += used;
        used = seq_table(w, w.ml, 2, (modes >> 2) & 3, p + at, n - at);
        if (used < 0) return false;
        at += used;
        if (at >= n) return false;
        BackBits b = back_open(p + at, n - at);
        if (!b.ok) return false;
        uint32_t sl = (uint32_t)back_read(b, w.ll.al), so = (uint32_t)back_read(b, w.of.al), sm = (uint32_t)back_read(b, w.ml.al);
        if (b.pos < 0) return false;
        for (int64_t i = 0; i < nseq; ++i) {
            const int oc = w.of.sym[so], lc = w.ll.sym[sl], mc = w.ml.sym[sm];
            if (oc > 31 || lc > 35 || mc > 52) return false;
            uint32_t lbase, mbase;
            int lnb, mnb;
            ll_code(lc, &lbase, &lnb);
            ml_code(mc, &mbase, &mnb);
            const uint64_t ov = (1ull << oc) + back_read(b, oc);
            const int64_t ml = (int64_t)mbase + (int64_t)back_read(b, mnb);
            const int64_t ll = (int64_t)lbase + (int64_t)back_read(b, lnb);
            if (b.pos < 0) return false;
            int64_t offset;
            if (ov > 3) {
                offset = (int64_t)(ov - 3);
                w.rep[2] = w.rep[1];
                w.rep[1] = w.rep[0];
                w.rep[0] = (uint32_t)offset;
            } else {
                int idx = (int)ov;
                if (ll != 0) idx--;
                if (idx == 0) {
                    offset = w.rep[0];
                } else {
                    offset = idx < 3 ? (int64_t)w.rep[idx] : (int64_t)w.rep[0] - 1;
                    if (offset <= 0) return false;
                    if (idx > 1) w.rep[2] = w.rep[1];
                    w.rep[1] = w.rep[0];
                    w.rep[0] = (uint32_t)offset;
                }
            }
            if (ll > n_lit - lp || ll + ml > cap - o) return false;
            memcpy(out + o, w.lit + lp, (size_t)ll);
            lp += ll;
            o += ll;
            if (offset > o) return false;
            const uint8_t* src = out + o - offset;
            if (offset >= ml) memcpy(out + o, src, (size_t)ml);
            else for (int64_t k = 0; k < ml; ++k) out[o + k] = src[k];
            o += ml;
            if (i + 1 < nseq) {  // state updates: literal length, match length, offset
                sl = w.ll.base[sl] + (uint32_t)back_read(b, w.ll.nb[sl]);
                sm = w.ml.base[sm] + (uint32_t)back_read(b, w.ml.nb[sm]);
                so = w.of.base[so] + (uint32_t)back_read(b, w.of.nb[so]);
                if (b.pos < 0) return false;
            }
        }
        if (b.pos != 0) return false;  // the bitstream is consumed exactly
    } else if (at != n) {
        return false;
    }
    const int64_t rest = n_lit - lp;
    if (rest > cap - o) return false;
    memcpy(out + o, w.lit + lp, (size_t)rest);
    *op = o + rest;
    return true;
}

// ---------------------------------------------------------------------------------------------- frame
// One zstd frame in[0..n) -> out[0..cap).  Returns the decoded size, or -1 (corrupt, unsupported, or does not fit).
M3D_HD inline int64_t decode_frame(Work& w, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap) {
    if (n < 6) return -1;
    const uint32_t magic = in[0] | (in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24);
    if (magic != MAGIC) return -1;
    const int fhd = in[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, dict_flag = fhd & 3;
    if (fhd & 0x08) return -1;  // reserved bit
    int64_t ip = 5;
    if (!single) ip += 1;  // window descriptor: the whole output buffer is the window here
    const int dict_bytes = dict_flag == 3 ? 4 : dict_flag;
    for (int i = 0; i < dict_bytes; ++i) {
        if (ip >= n) return -1;
        if (in[ip++] != 0) return -1;  // dictionaries are not supported
    }
    const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (ip + fcs_bytes > n) return -1;
    int64_t content = -1;
    if (fcs_bytes) {
        uint64_t v = 0;
        for (int i = 0; i < fcs_bytes; ++i) v |= (uint64_t)in[ip + i] << (8 * i);
        if (fcs_bytes == 2) v += 256;
        content = (int64_t)v;
        if (content < 0 || content > cap) return -1;
    }
    ip += fcs_bytes;
    w.huf_valid = 0;
    w.ll.valid = w.of.valid = w.ml.valid = 0;
    w.rep[0] = 1;
    w.rep[1] = 4;
    w.rep[2] = 8;
    int64_t op = 0;
    while (true) {
        if (ip + 3 > n) return -1;
        const uint32_t bh = in[ip] | (in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16);
        ip += 3;
        const int last = bh & 1, type = (bh >> 1) & 3;
        const int64_t size = bh >> 3;
        if (type == 0) {
            if (ip + size > n || size > cap - op) return -1;
            memcpy(out + op, in + ip, (size_t)size);
            ip += size;
            op += size;
        } else if (type == 1) {
            if (ip + 1 > n || size > cap - op) return -1;
            memset(out + op, in[ip], (size_t)size);
            ip += 1;
            op += size;
        } else if (type == 2) {
            if (size > MAX_BLOCK || ip + size > n) return -1;
            int64_t n_lit = 0;
            const int64_t lit_bytes = literals_decode(w, in + ip, size, &n_lit);
            if (lit_bytes < 0) return -1;
            if (!sequences_execute(w, in + ip + lit_bytes, size - lit_bytes, n_lit, out, cap, &op)) return -1;
            ip += size;
        } else {
            return -1;
        }
        if (last) break;
    }
    if (checksum) ip += 4;  // xxh64 of the content: not verified here
    if (ip > n) return -1;
    if (content >= 0 && op != content) return -1;
    return op;
}

}  // namespace m3d_zstd
