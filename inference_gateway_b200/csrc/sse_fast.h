// sse_fast.h -- whole-token shortcuts of the per-lane automaton (sse_kernel2.cu, v2_round): a member key `name":`, an
// integer literal and null / true / false are recognised from a 24-byte unaligned view of the payload and consumed in one
// step instead of one table transition per byte. Every shortcut is a STRICT SUBSET of what the table automaton accepts:
// it either reproduces exactly the state the byte-by-byte walk would reach (same p, st, slen, name id) or declines and
// leaves the lane untouched. The functions are pure (host + device) so that tests/test_fast_tokens_cpu.py can check that
// equivalence on the CPU against a table walk (tests/fast_tokens_check.cpp) -- there is no GPU where this is developed.
//
// Reference semantics behind the tables: encoding/json scanner.go (grammar), decode.go object() key match
// (exact name first, then case folding: a key of [a-z_] bytes can only match exactly because every struct tag of
// providers/types/common_types.go:271-478 is lower case).
#pragma once
#include <stdint.h>
#include <string.h>
#include "sse_tables.h"

#if defined(__CUDACC__)
#define SSE_HD __host__ __device__ __forceinline__
#else
#define SSE_HD inline
#endif

namespace ssefast {

constexpr int KH_SLOTS = 128;              // perfect hash over the struct-tag names (seed found at table build time)
constexpr int KH_MAXLEN = 21;              // name, closing quote and colon fit the 24-byte view
struct alignas(16) KeyEnt { uint32_t w[6]; uint32_t len; uint32_t name; };    // w: `name":` then zeros; name: id of sse_tables.h, 0xFF = empty slot
struct KeyHash { KeyEnt e[KH_SLOTS]; uint32_t seed; uint32_t pad[3]; };

SSE_HD uint32_t kh_slot(uint32_t w0, uint32_t n, uint32_t seed) { return ((w0 ^ (n * 0x9E3779B1u)) * seed) >> 25; }

// `names` / `ids`: the distinct names of the schema with their name ids. Returns 0 on success.
inline int build_keyhash(KeyHash &H, const char *const *names, const uint8_t *ids, int n_names) {
    uint8_t img[64][24];
    uint32_t len[64];
    if (n_names > 64) return -1;
    int m = 0;
    for (int i = 0; i < n_names; i++) {
        const size_t L = strlen(names[i]);
        if (L == 0 || L > (size_t)KH_MAXLEN) continue;                 // longer names take the table walk
        for (size_t k = 0; k < L; k++) if (!((names[i][k] >= 'a' && names[i][k] <= 'z') || names[i][k] == '_')) return -1;
        memset(img[m], 0, 24);
        memcpy(img[m], names[i], L);
        img[m][L] = '"'; if (L + 1 < 24) img[m][L + 1] = ':';
        len[m] = (uint32_t)L;
        img[m][23] = ids[i];                                            // parked here until the entry is written (L + 1 <= 22)
        m++;
    }
    uint32_t x = 0x2545F491u;
    for (int attempt = 0; attempt < 2000000; attempt++) {
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        const uint32_t seed = x | 1u;
        bool used[KH_SLOTS]; memset(used, 0, sizeof used);
        bool ok = true;
        for (int i = 0; i < m && ok; i++) {
            uint32_t w0; memcpy(&w0, img[i], 4);
            const uint32_t s = kh_slot(w0, len[i], seed);
            if (used[s]) ok = false; else used[s] = true;
        }
        if (!ok) continue;
        memset(&H, 0, sizeof H);
        for (int s = 0; s < KH_SLOTS; s++) H.e[s].name = 0xFFu;
        H.seed = seed;
        for (int i = 0; i < m; i++) {
            uint32_t w0; memcpy(&w0, img[i], 4);
            KeyEnt &E = H.e[kh_slot(w0, len[i], seed)];
            E.name = img[i][23]; E.len = len[i];
            uint8_t tmp[24]; memcpy(tmp, img[i], 24); tmp[23] = 0;
            memcpy(E.w, tmp, 24);
        }
        return 0;
    }
    return -1;
}

SSE_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) {          // bytes of hi:lo from bit s (s in {0, 8, 16, 24})
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}
SSE_HD uint32_t ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffsll((long long)v) - 1);
#else
    return (uint32_t)__builtin_ctzll(v);
#endif
}

// The 24 bytes from payload offset p: a = the aligned 16-byte window holding p, b = the window behind it (zeros when it
// lies outside the payload). Bytes past the two windows read as zero.
template <class V4>
SSE_HD void view24(const V4 &a, const V4 &b, uint32_t p, uint32_t (&v)[6]) {
    const uint32_t s = (p & 3u) * 8u;
    const bool k2 = (p & 8u) != 0, k1 = (p & 4u) != 0;
    const uint32_t x0 = k2 ? a.z : a.x, x1 = k2 ? a.w : a.y, x2 = k2 ? b.x : a.z, x3 = k2 ? b.y : a.w,
                   x4 = k2 ? b.z : b.x, x5 = k2 ? b.w : b.y, x6 = k2 ? 0u : b.z, x7 = k2 ? 0u : b.w;
    const uint32_t y0 = k1 ? x1 : x0, y1 = k1 ? x2 : x1, y2 = k1 ? x3 : x2, y3 = k1 ? x4 : x3,
                   y4 = k1 ? x5 : x4, y5 = k1 ? x6 : x5, y6 = k1 ? x7 : x6;
    v[0] = fsr(y0, y1, s); v[1] = fsr(y1, y2, s); v[2] = fsr(y2, y3, s);
    v[3] = fsr(y3, y4, s); v[4] = fsr(y4, y5, s); v[5] = fsr(y5, y6, s);
}
// how many of the view's bytes belong to the payload [p, pe) and were loaded
SSE_HD uint32_t view_avail(uint32_t p, uint32_t pe) {
    const uint32_t loaded = 32u - (p & 15u);
    uint32_t a = pe - p;
    if (a > loaded) a = loaded;
    return a > 24u ? 24u : a;
}
SSE_HD uint32_t byte_at(const uint32_t (&v)[6], uint32_t i) {         // i < 24
    const uint32_t j = i >> 2;
    const uint32_t a = (j & 1u) ? v[1] : v[0], b = (j & 1u) ? v[3] : v[2], c = (j & 1u) ? v[5] : v[4];
    const uint32_t w = (j & 4u) ? c : ((j & 2u) ? b : a);
    return (w >> ((i & 3u) * 8u)) & 0xFFu;
}
// index of the first byte whose flag (0x80 of its byte in m[]) is set; 24 when none. The flag words only have to be
// exact for the lowest flagged byte of each word.
SSE_HD uint32_t first_flag24(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t m4, uint32_t m5) {
    const uint64_t A = ((uint64_t)m1 << 32) | m0, B = ((uint64_t)m3 << 32) | m2, C = ((uint64_t)m5 << 32) | m4;
    if (A) return ctz64(A) >> 3;
    if (B) return 8u + (ctz64(B) >> 3);
    if (C) return 16u + (ctz64(C) >> 3);
    return 24u;
}
SSE_HD uint32_t eq_mask4(uint32_t w, uint32_t rep) { const uint32_t q = w ^ rep; return (q - 0x01010101u) & ~q & 0x80808080u; }
SSE_HD uint32_t nondigit_mask4(uint32_t w) {     // 0x80 in every byte that is not '0'..'9' (never misses one; exact for the lowest)
    return (~((w | 0x80808080u) - 0x30303030u) | (w + 0x46464646u) | w) & 0x80808080u;
}

// ---- a member key. The lane stands just behind the opening quote (state S_KSTR, no key byte consumed yet): v = view of
// the bytes from there, avail = view_avail. Returns the name id when the bytes are `name":` for a name of the schema --
// the walk would consume name, quote and colon (len + 2 bytes) and stand in S_VAL -- else 0xFFFFFFFF (decline).
// *len_out = length of the name.
SSE_HD uint32_t fast_key(const KeyHash &H, const uint32_t (&v)[6], uint32_t avail, uint32_t *len_out) {
    const uint32_t n = first_flag24(eq_mask4(v[0], 0x22222222u), eq_mask4(v[1], 0x22222222u), eq_mask4(v[2], 0x22222222u),
                                    eq_mask4(v[3], 0x22222222u), eq_mask4(v[4], 0x22222222u), eq_mask4(v[5], 0x22222222u));
    if (n == 0u || n > (uint32_t)KH_MAXLEN || n + 2u > avail) return 0xFFFFFFFFu;
    const KeyEnt &E = H.e[kh_slot(v[0], n, H.seed)];
    if (E.len != n || E.name == 0xFFu) return 0xFFFFFFFFu;
    const uint32_t d0 = v[0] ^ E.w[0], d1 = v[1] ^ E.w[1], d2 = v[2] ^ E.w[2], d3 = v[3] ^ E.w[3], d4 = v[4] ^ E.w[4], d5 = v[5] ^ E.w[5];
    const uint64_t A = ((uint64_t)d1 << 32) | d0, B = ((uint64_t)d3 << 32) | d2, C = ((uint64_t)d5 << 32) | d4;
    const uint32_t fd = A ? ctz64(A) >> 3 : (B ? 8u + (ctz64(B) >> 3) : (C ? 16u + (ctz64(C) >> 3) : 24u));    // first byte that differs from `name":`
    if (fd < n + 2u) return 0xFFFFFFFFu;
    *len_out = n;
    return E.name;
}

// ---- a value token at p (state S_VAL): v = view from p, avail = view_avail.
enum : uint32_t { FT_NONE = 0, FT_INT, FT_ZERO, FT_NULL, FT_TRUE, FT_FALSE };
// FT_INT: [1-9][0-9]{0,17} followed (inside the payload) by ',' '}' or ']' -- the walk would be in S_NINT with slen = len - 1
// when it meets the delimiter (A_NUM_END, delimiter looked up again). FT_ZERO: "0" + delimiter (S_NZERO). FT_NULL / FT_TRUE /
// FT_FALSE: the literal is complete inside the payload (the action fires on its last byte). *len_out = bytes of the token.
SSE_HD uint32_t fast_value(const uint32_t (&v)[6], uint32_t avail, uint32_t *len_out) {
    const uint32_t c0 = v[0] & 0xFFu;
    if (c0 - '0' <= 9u) {
        const uint32_t i = first_flag24(nondigit_mask4(v[0]), nondigit_mask4(v[1]), nondigit_mask4(v[2]),
                                        nondigit_mask4(v[3]), nondigit_mask4(v[4]), nondigit_mask4(v[5]));
        if (i > 18u || i >= avail) return FT_NONE;
        if (c0 == '0' && i != 1u) return FT_NONE;
        const uint32_t d = byte_at(v, i);
        if (d != ',' && d != '}' && d != ']') return FT_NONE;
        *len_out = i;
        return c0 == '0' ? FT_ZERO : FT_INT;
    }
    if (c0 == 'n') { if (avail >= 4u && v[0] == 0x6C6C756Eu) { *len_out = 4u; return FT_NULL; } return FT_NONE; }
    if (c0 == 't') { if (avail >= 4u && v[0] == 0x65757274u) { *len_out = 4u; return FT_TRUE; } return FT_NONE; }
    if (c0 == 'f') { if (avail >= 5u && v[0] == 0x736C6166u && (v[1] & 0xFFu) == 'e') { *len_out = 5u; return FT_FALSE; } return FT_NONE; }
    return FT_NONE;
}

// ---- the two shortcut phases of a round, shared verbatim by the decode kernel and the CPU equivalence check.
// LaneT: p, pe, win (.x .y .z .w: the aligned 16 bytes around p whenever p < pe), st, km, slen, sf, cur.
// Ops: ldwin(off) -> the aligned 16-byte window holding payload offset off; field(name) -> packed field of the member
// `name` of the object the lane is in (TTY_SKIP when skipping or unknown); number_end(end), lit_null(), lit_bool(),
// value_done(): the automaton's own actions (sse_kernel2.cu: v2_number_end, v2_null, A_LIT_TRUE / A_LIT_FALSE, value_done).
constexpr uint32_t STR_FLAGS = 15u;        // the per-string bits of Lane::sf (SF_STRMASK)

template <class LaneT>
SSE_HD uint32_t cur_byte(const LaneT &L) {
    const uint32_t wsel = (L.p >> 2) & 3u;
    const uint32_t w01 = (wsel & 1u) ? L.win.y : L.win.x, w23 = (wsel & 1u) ? L.win.w : L.win.z;
    const uint32_t w = (wsel & 2u) ? w23 : w01;
    return (w >> ((L.p & 3u) * 8u)) & 0xFFu;
}
template <class LaneT, class Ops>
SSE_HD void step1(LaneT &L, Ops &ops) {                                  // consume one byte
    L.p++;
    if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ops.ldwin(L.p);
}
template <class LaneT, class Ops, class V4>
SSE_HD void next_window(const LaneT &L, Ops &ops, V4 &r) {               // the window behind the one holding p; zeros outside the payload
    const uint32_t q = (L.p | 15u) + 1u;
    r.x = r.y = r.z = r.w = 0u;
    if (q < L.pe) r = ops.ldwin(q);
}
template <class LaneT, class Ops, class V4>
SSE_HD void lane_advance(LaneT &L, Ops &ops, uint32_t np, const V4 &w2) {   // p < np <= p + 24
    const uint32_t d = (np >> 4) - (L.p >> 4);
    L.p = np;
    if (d != 0u && np < L.pe) {
        if (d == 1u) L.win = w2;            // loaded: (old p | 15) + 1 <= np < pe
        else L.win = ops.ldwin(np);
    }
}

// REP: how many tokens (a key, or an integer / literal value) a lane may take in one call.
template <int REP, class LaneT, class Ops>
SSE_HD void fast_phases(const KeyHash &KH, LaneT &L, Ops &ops) {
    using namespace ssetab;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int rep = 0; rep < REP; rep++) {
    const uint32_t p_before = L.p;
    // phase K: [,] "name": ["]  -- a member key of the schema in one step (the table walk takes one transition per byte and
    // the A_KEY_END action). Entry: in front of the comma (S_AFTO), in front of the opening quote (S_KEY / S_OBJ0) or just
    // behind it (S_KSTR, nothing of the key consumed). The single-byte moves below are the table's own transitions.
    constexpr uint32_t KMASK = (1u << S_AFTO) | (1u << S_KEY) | (1u << S_OBJ0) | (1u << S_KSTR);
    static_assert(S_KSTR < 32, "state mask");
    if (L.p < L.pe && L.st < 32u && ((KMASK >> L.st) & 1u) && (L.st != S_KSTR || L.slen == 0u)) {
        uint32_t c = cur_byte(L);
        if (L.st == S_AFTO && c == ',') {                                   // tr[S_AFTO][','] = S_KEY
            L.st = S_KEY; L.km = TRIE_ROOT; L.sf &= ~STR_FLAGS; L.slen = 0;
            step1(L, ops);
            c = L.p < L.pe ? cur_byte(L) : 0u;
        }
        if ((L.st == S_KEY || L.st == S_OBJ0) && c == '"' && L.p < L.pe) {   // tr[S_KEY / S_OBJ0]['"'] = S_KSTR
            L.st = S_KSTR; L.km = TRIE_ROOT; L.sf &= ~STR_FLAGS; L.slen = 0;
            step1(L, ops);
        }
    }
    // one 24-byte view serves whichever token stands at p: a key (just behind its opening quote) or a value
    const bool key_pos = L.st == S_KSTR && L.slen == 0u && L.p < L.pe;
    bool val_pos = false;
    if (!key_pos && L.st == S_VAL && L.p < L.pe) { const uint32_t c = cur_byte(L); val_pos = c - '0' <= 9u || c == 'n' || c == 't' || c == 'f'; }
    if (key_pos || val_pos) {
        auto w2 = L.win;
        next_window(L, ops, w2);
        uint32_t v[6], n = 0;
        view24(L.win, w2, L.p, v);
        const uint32_t avail = view_avail(L.p, L.pe);
        if (key_pos) {
            const uint32_t name = fast_key(KH, v, avail, &n);
            if (name != 0xFFFFFFFFu) {                                       // = the key's bytes, A_KEY_END, tr[S_COLON][':'] = S_VAL
                L.cur = ops.field(name); L.st = S_VAL; L.km = TRIE_ROOT; L.slen = 0;
                lane_advance(L, ops, L.p + n + 2u, w2);
                if (L.p < L.pe && cur_byte(L) == '"') { L.st = S_VSTR; step1(L, ops); }   // tr[S_VAL]['"'] = S_VSTR
            }
        } else {
            // phase N: an integer in front of ',' '}' ']', or null / true / false, in one step
            const uint32_t kind = fast_value(v, avail, &n);
            if (kind != FT_NONE) {
                if (kind == FT_INT || kind == FT_ZERO) {                     // the walk stands on the delimiter: A_NUM_END, then the delimiter again
                    L.st = kind == FT_ZERO ? (uint32_t)S_NZERO : (uint32_t)S_NINT;
                    L.slen = n - 1u;
                    ops.number_end(L.p + n);
                } else if (kind == FT_NULL) ops.lit_null();
                else ops.lit_bool();
                ops.value_done();
                L.km = TRIE_ROOT; L.slen = 0;
                lane_advance(L, ops, L.p + n, w2);
            }
        }
    }
    if (L.p == p_before) break;            // nothing taken: the table walk goes on from here
    }
}

} // namespace ssefast
