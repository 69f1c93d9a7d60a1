// fn3_pickle.h — the reference's on-disk sample format, read without Python objects.
//
// findNeighbour3 stores every sample in GridFS as pickle.dumps(obj, protocol=2) of its compressed_sequence dict
// (src/mongoStore.py:344-363) and, at start-up, unpickles all of them into dicts of Python sets before persisting them
// again (src/findNeighbour3-server.py:358-362).  This is a minimal pickle virtual machine for exactly that object
// shape — {'A','C','G','T','N': set of int, 'M': {int: str}, 'invalid': 0} or {'invalid': 1} — that turns the bytes
// straight into the six sorted position lists the engine takes.  Anything else (a patch_and_consensus dict, numpy
// scalars, another protocol's opcodes) is reported as "not parsed": the caller unpickles those in Python.
// Host code only; no CUDA.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace fn3pickle {

enum { T_MARK, T_INT, T_STR, T_LIST, T_SETCLS, T_TUPLE1, T_SET, T_DICT, T_OTHER };

struct Val {
    uint8_t t = T_OTHER;
    int64_t i = 0;       // T_INT: value; T_STR: offset of the bytes in the blob; T_LIST/T_SET/T_TUPLE1/T_DICT: pool index
    int32_t n = 0;       // T_STR: length
};

struct Parsed {
    std::vector<int32_t> list[6];   // A,C,G,T,N,M (sorted)
    int invalid = 0;
};

struct Machine {
    const uint8_t* b; int64_t size; int64_t at = 0;
    std::vector<Val> stack, memo;
    std::vector<std::vector<int64_t>> lists;                        // list / set payloads (ints only)
    std::vector<std::vector<std::pair<Val, Val>>> dicts;
    bool ok = true;

    bool need(int64_t k) { if (at + k > size) { ok = false; return false; } return true; }
    uint32_t u32() { uint32_t v; memcpy(&v, b + at, 4); at += 4; return v; }
    void push(const Val& v) { stack.push_back(v); }
    bool pop(Val& v) { if (stack.empty()) { ok = false; return false; } v = stack.back(); stack.pop_back(); return true; }
    void put(uint32_t id) {
        if (stack.empty()) { ok = false; return; }
        if (id >= memo.size()) { if (id > (1u << 24) || (int64_t)id > size) { ok = false; return; } memo.resize(id + 1); }   // a memo id costs >= 2 bytes of pickle
        memo[id] = stack.back();
    }
    void get(uint32_t id) { if (id >= memo.size()) { ok = false; return; } push(memo[id]); }
    long long mark_pos() {
        for (long long k = (long long)stack.size() - 1; k >= 0; --k) if (stack[(size_t)k].t == T_MARK) return k;
        ok = false; return -1;
    }
    bool str_is(const Val& v, const char* s) const {
        const size_t n = strlen(s);
        return v.t == T_STR && (size_t)v.n == n && memcmp(b + v.i, s, n) == 0;
    }

    // runs to STOP; the result is the top of the stack
    bool run() {
        while (ok) {
            if (!need(1)) return false;
            const uint8_t op = b[at++];
            switch (op) {
            case 0x80: if (!need(1)) return false; if (b[at++] > 2) return false; break;          // PROTO <= 2
            case '}': { Val v; v.t = T_DICT; v.i = (int64_t)dicts.size(); dicts.emplace_back(); push(v); break; }
            case ']': { Val v; v.t = T_LIST; v.i = (int64_t)lists.size(); lists.emplace_back(); push(v); break; }
            case '(': {
                // fast path: MARK int int ... APPENDS right behind a list (a batch of up to 1000 set members)
                if (!stack.empty() && stack.back().t == T_LIST) {
                    std::vector<int64_t>& dst = lists[(size_t)stack.back().i];
                    const size_t n0 = dst.size();
                    int64_t p = at;
                    bool closed = false;
                    while (p < size) {
                        const uint8_t o = b[p];
                        if (o == 'K' && p + 2 <= size) { dst.push_back(b[p + 1]); p += 2; }
                        else if (o == 'M' && p + 3 <= size) { dst.push_back(b[p + 1] | (b[p + 2] << 8)); p += 3; }
                        else if (o == 'J' && p + 5 <= size) { int32_t x; memcpy(&x, b + p + 1, 4); dst.push_back(x); p += 5; }
                        else { closed = o == 'e'; break; }
                    }
                    if (closed) { at = p + 1; break; }
                    dst.resize(n0);                                  // something else in the batch: the general way
                }
                Val v; v.t = T_MARK; push(v); break;
            }
            case 'q': if (!need(1)) return false; put(b[at++]); break;                              // BINPUT
            case 'r': if (!need(4)) return false; put(u32()); break;                                // LONG_BINPUT
            case 'h': if (!need(1)) return false; get(b[at++]); break;                              // BINGET
            case 'j': if (!need(4)) return false; get(u32()); break;                                // LONG_BINGET
            case 'K': { if (!need(1)) return false; Val v; v.t = T_INT; v.i = b[at++]; push(v); break; }
            case 'M': { if (!need(2)) return false; Val v; v.t = T_INT; v.i = b[at] | (b[at + 1] << 8); at += 2; push(v); break; }
            case 'J': { if (!need(4)) return false; Val v; v.t = T_INT; v.i = (int32_t)u32(); push(v); break; }
            case 0x8a: {                                                                            // LONG1
                if (!need(1)) return false;
                const int n = b[at++];
                if (n > 8 || !need(n)) return false;
                int64_t x = 0;
                for (int k = 0; k < n; ++k) x |= (int64_t)b[at + k] << (8 * k);
                if (n && n < 8 && (b[at + n - 1] & 0x80)) x |= -((int64_t)1 << (8 * n));           // sign extension
                at += n;
                Val v; v.t = T_INT; v.i = x; push(v); break;
            }
            case 0x88: case 0x89: { Val v; v.t = T_INT; v.i = op == 0x88; push(v); break; }       // NEWTRUE / NEWFALSE
            case 'X': case 'T': {                                                                   // BINUNICODE / BINSTRING
                if (!need(4)) return false;
                const uint32_t n = u32();
                if (!need(n)) return false;
                Val v; v.t = T_STR; v.i = at; v.n = (int32_t)n; at += n; push(v); break;
            }
            case 'U': {                                                                             // SHORT_BINSTRING
                if (!need(1)) return false;
                const uint32_t n = b[at++];
                if (!need(n)) return false;
                Val v; v.t = T_STR; v.i = at; v.n = (int32_t)n; at += n; push(v); break;
            }
            case 'c': {                                                                             // GLOBAL module\nname\n
                const int64_t s0 = at;
                while (at < size && b[at] != '\n') ++at;
                const int64_t s1 = at++;
                const int64_t t0 = at;
                while (at < size && b[at] != '\n') ++at;
                const int64_t t1 = at++;
                if (at > size) return false;
                const bool builtin = (s1 - s0 == 11 && !memcmp(b + s0, "__builtin__", 11)) ||
                                     (s1 - s0 == 8 && !memcmp(b + s0, "builtins", 8));
                const bool is_set = t1 - t0 == 3 && !memcmp(b + t0, "set", 3);
                Val v; v.t = (builtin && is_set) ? T_SETCLS : T_OTHER; push(v);
                if (v.t == T_OTHER) return false;
                break;
            }
            case 'a': {                                                                             // APPEND
                Val x, l;
                if (!pop(x) || stack.empty()) return false;
                l = stack.back();
                if (l.t != T_LIST || x.t != T_INT) return false;
                lists[(size_t)l.i].push_back(x.i);
                break;
            }
            case 'e': {                                                                             // APPENDS
                const long long m = mark_pos();
                if (m < 1) return false;
                const Val l = stack[(size_t)m - 1];
                if (l.t != T_LIST) return false;
                std::vector<int64_t>& dst = lists[(size_t)l.i];
                for (size_t k = (size_t)m + 1; k < stack.size(); ++k) {
                    if (stack[k].t != T_INT) return false;
                    dst.push_back(stack[k].i);
                }
                stack.resize((size_t)m);
                break;
            }
            case 0x85: {                                                                            // TUPLE1
                Val x;
                if (!pop(x) || x.t != T_LIST) return false;
                x.t = T_TUPLE1; push(x); break;
            }
            case 'R': {                                                                             // REDUCE: set([..])
                Val arg, fn;
                if (!pop(arg) || !pop(fn)) return false;
                if (fn.t != T_SETCLS || arg.t != T_TUPLE1) return false;
                arg.t = T_SET; push(arg); break;
            }
            case 's': {                                                                             // SETITEM
                Val v, k;
                if (!pop(v) || !pop(k) || stack.empty()) return false;
                const Val d = stack.back();
                if (d.t != T_DICT) return false;
                dicts[(size_t)d.i].emplace_back(k, v);
                break;
            }
            case 'u': {                                                                             // SETITEMS
                const long long m = mark_pos();
                if (m < 1) return false;
                const Val d = stack[(size_t)m - 1];
                if (d.t != T_DICT || ((stack.size() - (size_t)m - 1) & 1)) return false;
                for (size_t k = (size_t)m + 1; k + 1 < stack.size(); k += 2) dicts[(size_t)d.i].emplace_back(stack[k], stack[k + 1]);
                stack.resize((size_t)m);
                break;
            }
            case '.': return ok && !stack.empty();                                                  // STOP
            default: return false;                                                                  // anything else: not ours
            }
        }
        return false;
    }
};

// ascending sort of non-negative int32 (set iteration order is hash order): LSD radix, 3 passes of 11 bits
inline void sort_positions(std::vector<int32_t>& v, std::vector<int32_t>& tmp) {
    const size_t n = v.size();
    if (n < 96) { std::sort(v.begin(), v.end()); return; }
    if (std::is_sorted(v.begin(), v.end())) return;
    tmp.resize(n);
    int32_t* a = v.data();
    int32_t* b = tmp.data();
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = 11 * pass;
        uint32_t cnt[2049];
        memset(cnt, 0, sizeof cnt);
        for (size_t i = 0; i < n; ++i) ++cnt[(((uint32_t)a[i] >> sh) & 2047u) + 1];
        for (int k = 0; k < 2048; ++k) cnt[k + 1] += cnt[k];
        for (size_t i = 0; i < n; ++i) b[cnt[((uint32_t)a[i] >> sh) & 2047u]++] = a[i];
        std::swap(a, b);
    }
    if (a != v.data()) memcpy(v.data(), a, n * sizeof(int32_t));      // three passes: the result sits in tmp
}

// 0 = parsed, valid sample; 1 = parsed, {'invalid': 1}; 2 = not a plain compressed_sequence pickle
inline int parse_sample(const uint8_t* blob, int64_t size, Parsed& out) {
    Machine m; m.b = blob; m.size = size;
    if (!m.run()) return 2;
    const Val top = m.stack.back();
    if (top.t != T_DICT) return 2;
    const auto& items = m.dicts[(size_t)top.i];
    static const char* names[6] = {"A", "C", "G", "T", "N", "M"};
    int seen[6] = {0, 0, 0, 0, 0, 0};
    int have_invalid = 0; int64_t invalid = 0;
    std::vector<int32_t> scratch;
    for (int l = 0; l < 6; ++l) out.list[l].clear();
    for (const auto& kv : items) {
        if (m.str_is(kv.first, "invalid")) {
            if (kv.second.t != T_INT) return 2;
            have_invalid = 1; invalid = kv.second.i;
            continue;
        }
        int l = -1;
        for (int c = 0; c < 6; ++c) if (m.str_is(kv.first, names[c])) l = c;
        if (l < 0 || seen[l]) return 2;
        seen[l] = 1;
        std::vector<int32_t>& dst = out.list[l];
        if (l < 5) {
            if (kv.second.t != T_SET) return 2;
            for (int64_t p : m.lists[(size_t)kv.second.i]) {
                if (p < 0 || p > 0x7fffffffll) return 2;
                dst.push_back((int32_t)p);
            }
        } else {
            if (kv.second.t != T_DICT) return 2;
            for (const auto& e : m.dicts[(size_t)kv.second.i]) {
                if (e.first.t != T_INT || e.second.t != T_STR || e.first.i < 0 || e.first.i > 0x7fffffffll) return 2;
                dst.push_back((int32_t)e.first.i);
            }
        }
        sort_positions(dst, scratch);
        if (std::adjacent_find(dst.begin(), dst.end()) != dst.end()) return 2;      // a set / dict has no duplicates
    }
    if (!have_invalid) return 2;
    const int n_seen = seen[0] + seen[1] + seen[2] + seen[3] + seen[4] + seen[5];
    if (invalid == 1 && n_seen == 0) { out.invalid = 1; return 1; }                  // {'invalid': 1}  (seqComparer.py:303)
    if (invalid != 0 || n_seen != 6) return 2;
    out.invalid = 0;
    return 0;
}

}  // namespace fn3pickle
