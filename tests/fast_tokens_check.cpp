// fast_tokens_check.cpp -- CPU equivalence check of the decode kernel's whole-token shortcuts (csrc/sse_fast.h).
//
// Test infrastructure only. The shortcut phases (ssefast::fast_phases: member keys, integers, null / true / false) are the
// SAME source the CUDA decode kernel compiles; here they run inside a host model of one lane of v2_round
// (csrc/sse_kernel2.cu): same 16-byte window discipline, same order of phases, the transition tables of sse_tables.h, and
// actions reduced to an event log (which key, which string span with which flags, which number span in which state, ...).
// Every payload is walked twice -- table transitions only, and with the shortcuts -- at all 16 alignments, with random
// bytes behind the payload's end; the two event logs and final states must be identical.
//
//   fast_tokens_check <names file> <corpus file: one payload per line> <fuzz rounds> <seed>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../inference_gateway_b200/csrc/sse_fast.h"

using namespace ssetab;
#ifndef FAST_REP
#define FAST_REP 1
#endif

struct V4 { uint32_t x, y, z, w; };
struct Ev { uint32_t kind, a, b, c; bool operator==(const Ev &o) const { return kind == o.kind && a == o.a && b == o.b && c == o.c; } };
enum { E_KEY = 1, E_VSTR, E_NUM, E_NULL, E_BOOL, E_OPEN, E_CLOSE, E_ELEM, E_COMMA_ARR, E_ERR };
constexpr uint32_t SF_ESC = 1, SF_HI = 2, SF_BAD = 8;

static DfaTables T;
static ssefast::KeyHash KH;
static uint32_t FINISH_NAME = 0xFFFF;
static uint64_t n_fast_keys = 0, n_fast_vals = 0, n_key_ends = 0;     // n_key_ends: A_KEY_END actions of the walks WITH shortcuts (declined keys)
static bool g_fast = false;

struct HLane {
    uint32_t p, pe; V4 win;
    uint32_t st, km, slen, sf, cur, depth;
    uint64_t ct[2];
    std::vector<Ev> log;
    const uint8_t *buf;
};

static V4 ldwin(const uint8_t *buf, uint32_t off) { V4 r; memcpy(&r, buf + (off & ~15u), 16); return r; }

static void value_done(HLane &L) {
    if (L.depth == 0) { L.st = S_END; return; }
    const uint32_t d = L.depth - 1;
    L.st = ((L.ct[d >> 6] >> (d & 63u)) & 1ull) ? (uint32_t)S_AFTA : (uint32_t)S_AFTO;
}
static void number_end(HLane &L, uint32_t end) { L.log.push_back({ E_NUM, end - L.slen - 1, end, L.st }); }

struct HOps {
    HLane &L;
    V4 ldwin(uint32_t off) const { return ::ldwin(L.buf, off); }
    uint32_t field(uint32_t name) const { n_fast_keys++; L.log.push_back({ E_KEY, name, 0, 0 }); return name; }
    void number_end(uint32_t end) const { n_fast_vals++; ::number_end(L, end); }
    void lit_null() const { n_fast_vals++; L.log.push_back({ E_NULL, L.p, 0, 0 }); }
    void lit_bool() const { n_fast_vals++; L.log.push_back({ E_BOOL, L.p, 0, 0 }); }
    void value_done() const { ::value_done(L); }
};

// returns true when the current byte has to be looked up again in the new state (v2_action)
static bool action(HLane &L, uint32_t t) {
    switch (t) {
    case A_OPEN_OBJ: case A_OPEN_ARR: {
        const bool arr = t == A_OPEN_ARR;
        if (L.depth >= 128) { L.log.push_back({ E_ERR, 1, 0, 0 }); L.p = L.pe - 1; L.st = S_END; return false; }
        L.ct[L.depth >> 6] = (L.ct[L.depth >> 6] & ~(1ull << (L.depth & 63))) | ((uint64_t)arr << (L.depth & 63));
        L.depth++;
        L.log.push_back({ E_OPEN, arr, L.cur, 0 });
        L.st = arr ? S_ARR0 : S_OBJ0;
        return false;
    }
    case A_CLOSE_OBJ: case A_CLOSE_ARR: L.depth--; L.log.push_back({ E_CLOSE, 0, 0, 0 }); value_done(L); return false;
    case A_KEY_END: {
        if (g_fast) n_key_ends++;
        const uint32_t name = (L.sf & (SF_ESC | SF_HI)) ? 0xFEu : (uint32_t)T.accept[L.km];
        // field(): a name the hash does not hold resolves to "unknown" on both paths; upper-case bytes (case folding) are the
        // table walk's business and never reach the shortcut
        L.cur = name;
        L.log.push_back({ E_KEY, name, 0, 0 });
        L.st = S_COLON;
        return false;
    }
    case A_VSTR_END:
        L.log.push_back({ E_VSTR, L.p - L.slen, L.slen, (L.sf & (SF_ESC | SF_HI | SF_BAD)) | (L.cur == FINISH_NAME ? (uint32_t)T.accept[L.km] << 8 : 0u) });
        value_done(L);
        return false;
    case A_BAD_STAY: L.sf |= SF_BAD; L.st = S_VSTR; return false;
    case A_BAD_REDO: L.sf |= SF_BAD; L.st = S_VSTR; return true;
    case A_NUM_END: number_end(L, L.p); value_done(L); return true;
    case A_LIT_TRUE: L.log.push_back({ E_BOOL, L.p - 3, 0, 0 }); value_done(L); return false;
    case A_LIT_FALSE: L.log.push_back({ E_BOOL, L.p - 4, 0, 0 }); value_done(L); return false;
    case A_LIT_NULL: L.log.push_back({ E_NULL, L.p - 3, 0, 0 }); value_done(L); return false;
    case A_ELEM_REDO: L.log.push_back({ E_ELEM, 0, 0, 0 }); L.cur = 0xAAAA; L.st = S_VAL; return true;
    case A_COMMA_ARR: L.log.push_back({ E_COMMA_ARR, 0, 0, 0 }); L.cur = 0xAAAA; L.st = S_VAL; return false;
    default: L.log.push_back({ E_ERR, 0, 0, 0 }); L.p = L.pe - 1; L.st = S_END; return false;
    }
}

static bool plain_string_byte(uint32_t c) { return !(c == '"' || c == '\\' || c < 0x20 || c >= 0x80); }

static void round(HLane &L, bool fast) {
    uint32_t pend = 0;
    if (fast) { HOps ops{ L }; ssefast::fast_phases<FAST_REP>(KH, L, ops); }
    // phase A: the string skip (up to 4 windows), bytewise here
    if (L.p < L.pe && L.st == S_VSTR && (L.km == TRIE_DEAD || (fast && L.cur != FINISH_NAME))) {
        uint32_t budget = 64;
        while (budget-- && L.p < L.pe && plain_string_byte(ssefast::cur_byte(L))) {
            L.p++; L.slen++;
            if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ldwin(L.buf, L.p);
        }
    }
    for (int k = 0; k < 2; k++) {
        if (L.p < L.pe && pend == 0) {
            const uint32_t c = ssefast::cur_byte(L);
            if (c != L.buf[L.p]) { fprintf(stderr, "window invariant broken at %u\n", L.p); exit(3); }
            const uint32_t e = T.clssym[c], cls = e & 63u;
            const bool in_str = L.st >= S_KSTR, in_tok = L.st >= S_NMINUS;
            const uint32_t t = T.tr[L.st * NCLS + cls];
            if (t < A_FIRST) {
                L.km = in_str ? (uint32_t)T.kt[L.km * NSYM + ((e >> 8) & 31u)] : (uint32_t)TRIE_ROOT;
                L.sf = in_str ? (L.sf | (e >> 13)) : (L.sf & ~15u);
                L.slen = in_tok ? L.slen + 1 : 0;
                L.st = t;
                L.p++;
                if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ldwin(L.buf, L.p);
            } else pend = t | (cls << 8) | (in_str ? 0x10000u : 0u) | (in_tok ? 0x20000u : 0u);
        }
    }
    if (pend) {
        uint32_t t = pend & 0xFFu;
        const uint32_t cls = (pend >> 8) & 0xFFu;
        for (;;) {
            if (!action(L, t)) break;
            t = T.tr[L.st * NCLS + cls];
            if (t < A_FIRST) { L.st = t; break; }
        }
        const bool in_str = (pend & 0x10000u) != 0;
        L.km = in_str ? (uint32_t)TRIE_DEAD : (uint32_t)TRIE_ROOT;
        const uint32_t nf = (cls == C_BSLASH ? SF_ESC : 0u) | (cls >= C_H80 ? SF_HI : 0u);
        L.sf = in_str ? (L.sf | nf) : (L.sf & ~15u);
        L.slen = (pend & 0x20000u) ? L.slen + 1 : 0;
        L.p++;
        if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ldwin(L.buf, L.p);
    }
}

static void walk(HLane &L, const uint8_t *buf, uint32_t ps, uint32_t pe, bool fast) {
    L.buf = buf; L.p = ps; L.pe = pe; L.st = S_VAL; L.km = TRIE_ROOT; L.slen = 0; L.sf = 0; L.cur = 0xBBBB; L.depth = 0; L.ct[0] = L.ct[1] = 0;
    L.log.clear();
    g_fast = fast;
    memset(&L.win, 0, sizeof L.win);
    if (L.p < L.pe) L.win = ldwin(buf, L.p);
    uint32_t guard = 0;
    while (L.p < L.pe) { round(L, fast); if (++guard > 4u * (pe - ps) + 16u) { fprintf(stderr, "no progress\n"); exit(3); } }
    // v2_finish_line: a number may end with the payload
    if (L.depth == 0 && (L.st == S_NZERO || L.st == S_NINT || L.st == S_NFRAC || L.st == S_NEXP)) { number_end(L, L.pe); L.st = S_END; }
    L.log.push_back({ 100, L.st, L.depth, L.p });
}

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

static int check(const std::string &doc, const char *what) {
    static std::vector<uint8_t> buf;
    HLane A, B;
    for (uint32_t al = 0; al < 16; al++) {
        const uint32_t ps = 32 + al, pe = ps + (uint32_t)doc.size();
        buf.assign(pe + 64 + 16, 0);
        for (size_t i = 0; i < buf.size(); i++) { static const char junk[] = "\"\\,:{}[]0123456789nulltruefalse id"; buf[i] = (uint8_t)junk[rnd() % (sizeof junk - 1)]; }
        memcpy(buf.data() + ps, doc.data(), doc.size());
        const uint8_t *base = buf.data();
        // 16-byte aligned base as in the arena
        std::vector<uint8_t> al_buf(buf.size() + 16);
        uint8_t *q = al_buf.data() + ((16 - ((uintptr_t)al_buf.data() & 15)) & 15);
        memcpy(q, base, buf.size());
        walk(A, q, ps, pe, false);
        walk(B, q, ps, pe, true);
        if (!(A.log == B.log)) {
            fprintf(stderr, "MISMATCH (%s) alignment %u, %zu vs %zu events, doc: %.*s\n", what, al, A.log.size(), B.log.size(), (int)doc.size(), doc.c_str());
            for (size_t i = 0; i < A.log.size() || i < B.log.size(); i++) {
                Ev a = i < A.log.size() ? A.log[i] : Ev{ 0, 0, 0, 0 }, b = i < B.log.size() ? B.log[i] : Ev{ 0, 0, 0, 0 };
                fprintf(stderr, "  %c [%zu] table %u:%u,%u,%u   fast %u:%u,%u,%u\n", a == b ? ' ' : '!', i, a.kind, a.a, a.b, a.c, b.kind, b.a, b.b, b.c);
            }
            return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage\n"); return 2; }
    std::vector<std::string> names;
    { FILE *f = fopen(argv[1], "r"); if (!f) return 2; char line[256]; while (fgets(line, sizeof line, f)) { std::string s(line); while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); if (!s.empty()) names.push_back(s); } fclose(f); }
    std::vector<FieldSrc> fs;
    for (size_t i = 0; i < names.size(); i++) fs.push_back({ (uint8_t)(1 + i % 11), names[i].c_str(), TTY_STR, 0, 0 });
    static const char *fin_names[] = { "stop", "tool_calls", "length", "content_filter", "function_call" };
    static const uint8_t fin_vals[] = { 1, 2, 3, 4, 5 };
    const char *nm[NNAMES]; int n_names = 0;
    if (build_tables(T, fs.data(), (int)fs.size(), fin_names, fin_vals, 5, nm, &n_names) != 0) { fprintf(stderr, "build_tables failed\n"); return 2; }
    uint8_t ids[NNAMES]; for (int i = 0; i < NNAMES; i++) ids[i] = (uint8_t)i;
    if (ssefast::build_keyhash(KH, nm, ids, n_names) != 0) { fprintf(stderr, "build_keyhash failed\n"); return 2; }
    for (int i = 0; i < n_names; i++) if (!strcmp(nm[i], "finish_reason")) FINISH_NAME = (uint32_t)i;
    // every name of at most KH_MAXLEN bytes must be in the hash under its own id
    for (int i = 0; i < n_names; i++) {
        const size_t L = strlen(nm[i]);
        if (L > (size_t)ssefast::KH_MAXLEN) continue;
        uint8_t img[24] = { 0 }; memcpy(img, nm[i], L); img[L] = '"'; img[L + 1] = ':';
        uint32_t v[6]; memcpy(v, img, 24);
        uint32_t n = 0;
        if (ssefast::fast_key(KH, v, 24, &n) != (uint32_t)i || n != L) { fprintf(stderr, "hash misses %s\n", nm[i]); return 1; }
    }
    std::vector<std::string> corpus;
    { FILE *f = fopen(argv[2], "rb"); if (!f) return 2; std::string cur; int ch; while ((ch = fgetc(f)) != EOF) { if (ch == '\n') { if (!cur.empty()) corpus.push_back(cur); cur.clear(); } else cur.push_back((char)ch); } if (!cur.empty()) corpus.push_back(cur); fclose(f); }
    const long fuzz = atol(argv[3]);
    rng_state ^= (uint64_t)atol(argv[4]) * 0x9E3779B97F4A7C15ull;
    for (const std::string &d : corpus) if (check(d, "corpus")) return 1;
    const uint64_t clean_keys = n_fast_keys, clean_key_ends = n_key_ends, clean_vals = n_fast_vals;
    static const char alpha[] = "\"\\,:{}[]0123456789-+.eEntfalsru _ABCxyz\t\n/";
    for (long it = 0; it < fuzz; it++) {
        std::string d = corpus[rnd() % corpus.size()];
        if (d.size() > 600) d.resize(600 - rnd() % 200);                     // long content adds nothing here
        const uint32_t nm_ = 1 + rnd() % 3;
        for (uint32_t k = 0; k < nm_ && !d.empty(); k++) {
            const uint32_t pos = rnd() % d.size();
            switch (rnd() % 6) {
            case 0: d[pos] = alpha[rnd() % (sizeof alpha - 1)]; break;
            case 1: d.erase(pos, 1 + rnd() % 3); break;
            case 2: d.insert(pos, 1, alpha[rnd() % (sizeof alpha - 1)]); break;
            case 3: d.resize(pos); break;                                      // truncation: every cut position matters for the bounds
            case 4: d[pos] = (char)(rnd() & 0xFF); break;
            default: { const uint32_t a = rnd() % d.size(), n = 1 + rnd() % 24; d.insert(pos, d.substr(a, n)); break; }
            }
        }
        if (check(d, "fuzz")) return 1;
    }
    printf("{\"docs\": %zu, \"fuzz\": %ld, \"clean_fast_keys\": %llu, \"clean_table_key_ends\": %llu, \"clean_fast_values\": %llu, \"fast_keys\": %llu, \"fast_values\": %llu}\n",
           corpus.size(), fuzz, (unsigned long long)clean_keys, (unsigned long long)clean_key_ends, (unsigned long long)clean_vals,
           (unsigned long long)n_fast_keys, (unsigned long long)n_fast_vals);
    return 0;
}
