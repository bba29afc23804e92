// zstd_lanes.cuh -- the zstd frame decoder of zstd_decode.cuh re-arranged for a team of lanes (a warp).
//
// What the frames of a readout chunk contain decides the arrangement (tools/zstd_frame_stats.py): per 256 KiB Blosc
// block ~115 KB of Huffman literals in FOUR independent streams and a few hundred sequences.  So: lane 0 parses the
// literals header and builds the Huffman table, lanes 0..3 decode one stream each, every lane runs the (short) FSE
// sequence chain redundantly -- it is uniform, all lanes read the same bytes -- and the literal and match copies are
// shared by the team exactly as the LZ4 decoder shares them.
//
// STATUS: the logic is pinned on the HOST through the one-lane policy (`m3d_zstd_decode_builtin_lanes`,
// tests/test_cpu_zarr_store.py: same frames as the serial decoder); the warp policy / `blosc_zstd_decode_kernel_v2`
// (M3D_ZARR_GPU_ZSTD=2) is bit-exact on the B200 against the host decode (tests/test_gpu_zarr_store.py, both modes, plus a
// damaged frame).  Measurements: profiles/r2_zstd_device.txt.
#pragma once
#include "zstd_decode.cuh"

namespace m3d_zstd {

struct OneLane {
    static M3D_HD inline int lane() { return 0; }
    static M3D_HD inline int width() { return 1; }
    static M3D_HD inline void sync() {}
    static M3D_HD inline int all(int ok) { return ok; }
    static M3D_HD inline int64_t bcast(int64_t v) { return v; }
    static M3D_HD inline void copy(uint8_t* d, const uint8_t* s, int64_t n) { memcpy(d, s, (size_t)n); }
    static M3D_HD inline void fill(uint8_t* d, int v, int64_t n) { memset(d, v, (size_t)n); }
    static M3D_HD inline void match(uint8_t* d, int64_t off, int64_t n) {
        if (off >= n) memcpy(d, d - off, (size_t)n);
        else for (int64_t i = 0; i < n; ++i) d[i] = d[i - off];
    }
};

#ifdef __CUDACC__
struct Warp32 {
    static __device__ inline int lane() { return threadIdx.x & 31; }
    static __device__ inline int width() { return 32; }
    static __device__ inline void sync() { __syncwarp(); }
    static __device__ inline int all(int ok) { return __all_sync(0xffffffffu, ok); }
    static __device__ inline int64_t bcast(int64_t v) {
        const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)(uint64_t)v, 0);
        const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)v >> 32), 0);
        return (int64_t)(((uint64_t)hi << 32) | lo);
    }
    static __device__ inline void copy(uint8_t* d, const uint8_t* s, int64_t n) {
        const int lane = threadIdx.x & 31;
        if ((((uintptr_t)d ^ (uintptr_t)s) & 3u) == 0 && n >= 64) {  // same alignment: 32-bit words
            int64_t head = (4 - (int64_t)((uintptr_t)d & 3u)) & 3;
            if (lane < head) d[lane] = s[lane];
            const int64_t nw = (n - head) >> 2;
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(s + head);
            uint32_t* dw = reinterpret_cast<uint32_t*>(d + head);
            int64_t i = lane;
            for (; i + 96 < nw; i += 128) {  // four independent loads in flight per lane
                const uint32_t a = sw[i], b = sw[i + 32], c = sw[i + 64], e = sw[i + 96];
                dw[i] = a;
                dw[i + 32] = b;
                dw[i + 64] = c;
                dw[i + 96] = e;
            }
            for (; i < nw; i += 32) dw[i] = sw[i];
            const int64_t done = head + (nw << 2);
            if (lane < n - done) d[done + lane] = s[done + lane];
            return;
        }
        int64_t i = lane;
        for (; i + 96 < n; i += 128) {
            const uint8_t a = s[i], b = s[i + 32], c = s[i + 64], e = s[i + 96];
            d[i] = a;
            d[i + 32] = b;
            d[i + 64] = c;
            d[i + 96] = e;
        }
        for (; i < n; i += 32) d[i] = s[i];
    }
    static __device__ inline void fill(uint8_t* d, int v, int64_t n) {
        const int lane = threadIdx.x & 31;
        if (n >= 64) {
            int64_t head = (4 - (int64_t)((uintptr_t)d & 3u)) & 3;
            if (lane < head) d[lane] = (uint8_t)v;
            const int64_t nw = (n - head) >> 2;
            uint32_t* dw = reinterpret_cast<uint32_t*>(d + head);
            const uint32_t vv = (uint32_t)(uint8_t)v * 0x01010101u;
            for (int64_t i = lane; i < nw; i += 32) dw[i] = vv;
            const int64_t done = head + (nw << 2);
            if (lane < n - done) d[done + lane] = (uint8_t)v;
            return;
        }
        for (int64_t i = lane; i < n; i += 32) d[i] = (uint8_t)v;
    }
    static __device__ inline void match(uint8_t* d, int64_t off, int64_t n) {
        const uint8_t* src = d - off;
        const int lane = threadIdx.x & 31;
        if (off >= n) {
            copy(d, src, n);
        } else if (off == 1) {  // a run of one byte (all-zero bit rows of the shuffled high bytes)
            fill(d, src[0], n);
        } else if (off >= 32) {
            // overlapping, long period: waves of `off` bytes; a wave only reads what earlier waves (or the bytes
            // before the match) wrote
            for (int64_t base = 0; base < n; base += off) {
                const int64_t len = off < n - base ? off : n - base;
                for (int64_t i = lane; i < len; i += 32) d[base + i] = d[base + i - off];
                __syncwarp();
            }
        } else {  // short period: byte i is byte (i mod offset) of the bytes before the match
            const uint32_t o = (uint32_t)off, step = 32u % o;
            uint32_t phase = (uint32_t)lane % o;
            for (int64_t i = lane; i < n; i += 32) {
                d[i] = src[phase];
                phase += step;
                if (phase >= o) phase -= o;
            }
        }
    }
};
#endif

// where the literals of a compressed block come from (offsets relative to the start of the literals section)
struct LitPlan {
    int kind;  // 0 raw, 1 RLE, 2 Huffman
    int n_streams;
    int rle_byte;
    int64_t regen, section_bytes, raw_off;
    int64_t src_off[4], src_len[4], out_off[4], out_n[4];
};

// One lane: parse the literals section header at p[0..n), build the Huffman table when the block carries one, and
// say where each stream is.  Same rules as literals_decode.
M3D_HD inline bool literals_prepare(Work& w, const uint8_t* p, int64_t n, LitPlan& plan) {
    if (n < 1) return false;
    const int type = p[0] & 3, fmt = (p[0] >> 2) & 3;
    if (type < 2) {
        int64_t hdr, regen;
        if ((fmt & 1) == 0) {
            hdr = 1;
            regen = p[0] >> 3;
        } else if (fmt == 1) {
            if (n < 2) return false;
            hdr = 2;
            regen = (p[0] >> 4) | ((int64_t)p[1] << 4);
        } else {
            if (n < 3) return false;
            hdr = 3;
            regen = (p[0] >> 4) | ((int64_t)p[1] << 4) | ((int64_t)p[2] << 12);
        }
        if (regen > MAX_BLOCK) return false;
        plan.kind = type;
        plan.n_streams = 0;
        plan.regen = regen;
        plan.raw_off = hdr;
        if (type == 0) {
            if (hdr + regen > n) return false;
            plan.section_bytes = hdr + regen;
        } else {
            if (hdr + 1 > n) return false;
            plan.rle_byte = p[hdr];
            plan.section_bytes = hdr + 1;
        }
        return true;
    }
    int64_t hdr, regen, comp;
    int streams;
    if (fmt == 0 || fmt == 1) {
        if (n < 3) return false;
        const uint32_t v = p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16);
        hdr = 3;
        streams = fmt == 0 ? 1 : 4;
        regen = (v >> 4) & 0x3FF;
        comp = (v >> 14) & 0x3FF;
    } else if (fmt == 2) {
        if (n < 4) return false;
        const uint32_t v = p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        hdr = 4;
        streams = 4;
        regen = (v >> 4) & 0x3FFF;
        comp = (v >> 18) & 0x3FFF;
    } else {
        if (n < 5) return false;
        const uint64_t v = (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24) |
                           ((uint64_t)p[4] << 32);
        hdr = 5;
        streams = 4;
        regen = (int64_t)((v >> 4) & 0x3FFFF);
        comp = (int64_t)((v >> 22) & 0x3FFFF);
    }
    if (regen > MAX_BLOCK || hdr + comp > n) return false;
    int64_t at = hdr, left = comp;
    if (type == 2) {
        const int64_t tree = huf_read_tree(w, p + at, left);
        if (tree < 0) return false;
        at += tree;
        left -= tree;
    } else if (!w.huf_valid) {
        return false;
    }
    plan.kind = 2;
    plan.n_streams = streams;
    plan.regen = regen;
    plan.section_bytes = hdr + comp;
    if (streams == 1) {
        plan.src_off[0] = at;
        plan.src_len[0] = left;
        plan.out_off[0] = 0;
        plan.out_n[0] = regen;
        return true;
    }
    if (left < 6) return false;
    const uint8_t* q = p + at;
    const int64_t s1 = q[0] | (q[1] << 8), s2 = q[2] | (q[3] << 8), s3 = q[4] | (q[5] << 8);
    const int64_t s4 = left - 6 - s1 - s2 - s3;
    const int64_t each = (regen + 3) / 4, last = regen - 3 * each;
    if (s4 < 1 || last < 0) return false;
    const int64_t len[4] = {s1, s2, s3, s4};
    int64_t src = at + 6;
    for (int s = 0; s < 4; ++s) {
        plan.src_off[s] = src;
        plan.src_len[s] = len[s];
        plan.out_off[s] = s * each;
        plan.out_n[s] = s < 3 ? each : last;
        src += len[s];
    }
    return true;
}

// Sequences section, team version of sequences_execute: lane 0 builds the three code tables, every lane walks the
// bitstream (uniform), the copies are shared.
template <class L>
M3D_HD inline bool sequences_lanes(Work& w, const uint8_t* p, int64_t n, int64_t n_lit, uint8_t* out, int64_t cap, int64_t* op) {
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
    uint32_t rep[3] = {w.rep[0], w.rep[1], w.rep[2]};
    if (nseq > 0) {
        if (at >= n) return false;
        const int modes = p[at++];
        if (modes & 3) return false;
        int64_t after = -1;
        if (L::lane() == 0) {
            int64_t a = at;
            int64_t used = seq_table(w, w.ll, 0, (modes >> 6) & 3, p + a, n - a);
            if (used >= 0) {
                a += used;
                used = seq_table(w, w.of, 1, (modes >> 4) & 3, p + a, n - a);
            }
            if (used >= 0) {
                a += used;
                used = seq_table(w, w.ml, 2, (modes >> 2) & 3, p + a, n - a);
            }
            if (used >= 0) after = a + used;
        }
        L::sync();  // the tables are visible to the team
        after = L::bcast(after);
        if (after < 0 || after >= n) return false;
        at = after;
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
                rep[2] = rep[1];
                rep[1] = rep[0];
                rep[0] = (uint32_t)offset;
            } else {
                int idx = (int)ov;
                if (ll != 0) idx--;
                if (idx == 0) {
                    offset = rep[0];
                } else {
                    offset = idx < 3 ? (int64_t)rep[idx] : (int64_t)rep[0] - 1;
                    if (offset <= 0) return false;
                    if (idx > 1) rep[2] = rep[1];
                    rep[1] = rep[0];
                    rep[0] = (uint32_t)offset;
                }
            }
            if (ll > n_lit - lp || ll + ml > cap - o) return false;
            L::copy(out + o, w.lit + lp, ll);
            lp += ll;
            o += ll;
            if (offset > o) return false;
            L::sync();  // these literals and the previous match are visible before the match reads them
            L::match(out + o, offset, ml);
            o += ml;
            if (i + 1 < nseq) {
                sl = w.ll.base[sl] + (uint32_t)back_read(b, w.ll.nb[sl]);
                sm = w.ml.base[sm] + (uint32_t)back_read(b, w.ml.nb[sm]);
                so = w.of.base[so] + (uint32_t)back_read(b, w.of.nb[so]);
                if (b.pos < 0) return false;
            }
        }
        if (b.pos != 0) return false;
    } else if (at != n) {
        return false;
    }
    const int64_t rest = n_lit - lp;
    if (rest > cap - o) return false;
    L::copy(out + o, w.lit + lp, rest);
    *op = o + rest;
    L::sync();  // every lane has read w.rep and the tables; the output so far is visible
    if (L::lane() == 0) {
        w.rep[0] = rep[0];
        w.rep[1] = rep[1];
        w.rep[2] = rep[2];
    }
    L::sync();
    return true;
}

// One zstd frame, decoded by a team.  Every lane calls it with the same arguments; every lane gets the same result.
template <class L>
M3D_HD inline int64_t decode_frame_lanes(Work& w, LitPlan& plan, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap) {
    if (n < 6) return -1;
    const uint32_t magic = in[0] | (in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24);
    if (magic != MAGIC) return -1;
    const int fhd = in[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, dict_flag = fhd & 3;
    if (fhd & 0x08) return -1;
    int64_t ip = 5;
    if (!single) ip += 1;
    const int dict_bytes = dict_flag == 3 ? 4 : dict_flag;
    for (int i = 0; i < dict_bytes; ++i) {
        if (ip >= n) return -1;
        if (in[ip++] != 0) return -1;
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
    if (L::lane() == 0) {
        w.huf_valid = 0;
        w.ll.valid = w.of.valid = w.ml.valid = 0;
        w.rep[0] = 1;
        w.rep[1] = 4;
        w.rep[2] = 8;
    }
    L::sync();
    int64_t op = 0;
    while (true) {
        if (ip + 3 > n) return -1;
        const uint32_t bh = in[ip] | (in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16);
        ip += 3;
        const int last = bh & 1, type = (bh >> 1) & 3;
        const int64_t size = bh >> 3;
        if (type == 0) {
            if (ip + size > n || size > cap - op) return -1;
            L::copy(out + op, in + ip, size);
            L::sync();
            ip += size;
            op += size;
        } else if (type == 1) {
            if (ip + 1 > n || size > cap - op) return -1;
            L::fill(out + op, in[ip], size);
            L::sync();
            ip += 1;
            op += size;
        } else if (type == 2) {
            if (size > MAX_BLOCK || ip + size > n) return -1;
            const uint8_t* p = in + ip;
            int ok = 1;
            if (L::lane() == 0) ok = literals_prepare(w, p, size, plan) ? 1 : 0;
            L::sync();  // the plan and the Huffman table are visible to the team
            if (!L::bcast(ok)) return -1;
            if (plan.kind == 0) {
                L::copy(w.lit, p + plan.raw_off, plan.regen);
            } else if (plan.kind == 1) {
                L::fill(w.lit, plan.rle_byte, plan.regen);
            } else {
                int good = 1;
                for (int s = L::lane(); s < plan.n_streams; s += L::width())  // one lane per Huffman stream
                    good &= huf_decode_stream(w, p + plan.src_off[s], plan.src_len[s], w.lit + plan.out_off[s], plan.out_n[s]) ? 1 : 0;
                if (!L::all(good)) return -1;
            }
            L::sync();  // the literals are visible
            const int64_t lit_bytes = plan.section_bytes, n_lit = plan.regen;
            if (!sequences_lanes<L>(w, p + lit_bytes, size - lit_bytes, n_lit, out, cap, &op)) return -1;
            ip += size;
        } else {
            return -1;
        }
        if (last) break;
    }
    if (checksum) ip += 4;
    if (ip > n) return -1;
    if (content >= 0 && op != content) return -1;
    return op;
}

}  // namespace m3d_zstd
