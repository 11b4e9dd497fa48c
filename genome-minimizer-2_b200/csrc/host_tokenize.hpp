// host_tokenize.hpp — part of libgm2.so (included by gm2.cu; one translation unit).  HOST ONLY.
// gm2_tokenize_pickle: the gene-name lists container -> id CSR without creating Python objects.
//
// The reference reads its lists with `np.load(genes_path, allow_pickle=True).tolist()`
// (minimizer_2.py:456, :518); the file is what `np.save(..., allow_pickle=True)` wrote at
// binary_converter.py:71 / :117: a .npy header followed by ONE pickle of an object ndarray.  pickle
// memoises every str object, so after the first ~V names the stream is a sequence of memo
// references (BINGET / LONG_BINGET) — already a token stream.  This is a pickle virtual machine
// restricted to what such a file can contain: it resolves each str to its vocabulary id once
// (at the opcode that defines it) and copies ids afterwards.  Anything outside the subset
// (text-mode opcodes, Python-2 strings, persistent ids, out-of-band buffers, list items that are
// not str) returns GM2_ERR_UNSUPPORTED and the caller falls back to NumPy's own loader, so the
// accepted inputs are exactly those for which `name in list` is plain str equality
// (minimizer_2.py:62).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace gm2tok {

// Every pickle value the machine tracks is ONE int32 ("item"), on the stack, in the memo and inside
// sequences alike, so that a memo reference is a 4-byte copy and APPENDS is a memcpy:
//   >= 0  a str that is vocabulary entry `item`          -1  a str outside the vocabulary
//   -2    any other object (int, bytes, None, dict, numpy scalar, ...)
//   <= -3 a list or tuple: index -3 - item into `seqs`
typedef int32_t Item;
static const Item IT_UNKNOWN_STR = -1, IT_OTHER = -2;
static inline bool is_seq(Item x) { return x <= -3; }
static inline int32_t seq_index(Item x) { return -3 - x; }

struct Machine {
    const uint8_t* p; const uint8_t* end;
    std::vector<Item> stack, memo;
    std::vector<size_t> marks;                      // stack heights at each open MARK (pickle's metastack)
    std::vector<std::vector<Item>> seqs;
    int32_t last_state = -1;                        // sequence index of the state of the last BUILD
    std::unordered_map<std::string_view, int32_t> vocab;

    bool need(size_t n) const { return (size_t)(end - p) >= n; }
    template <typename T> T rd() { T v; memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
    Item new_seq() { seqs.emplace_back(); return -3 - ((int32_t)seqs.size() - 1); }
    // closes the topmost MARK: returns the stack height it was opened at (items above it belong to
    // the opcode being executed), or -1 when there is none
    int64_t take_mark() {
        if (marks.empty() || marks.back() > stack.size()) return -1;
        const size_t m = marks.back(); marks.pop_back();
        return (int64_t)m;
    }
    bool push_str(uint64_t n) {
        if (!need(n)) return false;
        auto it = vocab.find(std::string_view(reinterpret_cast<const char*>(p), (size_t)n));
        stack.push_back(it == vocab.end() ? IT_UNKNOWN_STR : it->second);
        p += n;
        return true;
    }
    bool skip_push_other(uint64_t n) {
        if (!need(n)) return false;
        p += n; stack.push_back(IT_OTHER);
        return true;
    }
    bool pop(size_t n) { if (stack.size() < n) return false; stack.resize(stack.size() - n); return true; }
    bool seq_from_top(size_t n) {                    // TUPLE1..3, and LIST / TUPLE once the mark is resolved
        if (stack.size() < n) return false;
        const Item s = new_seq();
        seqs[seq_index(s)].assign(stack.end() - (ptrdiff_t)n, stack.end());
        stack.resize(stack.size() - n);
        stack.push_back(s);
        return true;
    }

    // 0 ok (STOP reached), 1 unsupported, 2 corrupt
    int run() {
        while (p < end) {
            const uint8_t op = *p++;
            switch (op) {
            case 0x80: if (!need(1)) return 2; p += 1; break;                                 // PROTO
            case 0x95: if (!need(8)) return 2; p += 8; break;                                 // FRAME
            case '.': return 0;                                                               // STOP
            case '(': marks.push_back(stack.size()); break;                                   // MARK
            case '0':                                                                         // POP
                if (!marks.empty() && marks.back() == stack.size()) marks.pop_back();
                else if (!pop(1)) return 2;
                break;
            case '1': { const int64_t m = take_mark(); if (m < 0) return 2; stack.resize((size_t)m); break; }   // POP_MARK
            case '2': if (stack.empty()) return 2; stack.push_back(stack.back()); break;      // DUP
            case 'N': case 0x88: case 0x89: stack.push_back(IT_OTHER); break;             // NONE NEWTRUE NEWFALSE
            case 'J': if (!skip_push_other(4)) return 2; break;                               // BININT
            case 'K': if (!skip_push_other(1)) return 2; break;                               // BININT1
            case 'M': if (!skip_push_other(2)) return 2; break;                               // BININT2
            case 'G': if (!skip_push_other(8)) return 2; break;                               // BINFLOAT
            case 0x8a: { if (!need(1)) return 2; const uint8_t n = rd<uint8_t>(); if (!skip_push_other(n)) return 2; break; }     // LONG1
            case 0x8b: { if (!need(4)) return 2; const uint32_t n = rd<uint32_t>(); if (!skip_push_other(n)) return 2; break; }   // LONG4
            case 0x8c: { if (!need(1)) return 2; const uint8_t n = rd<uint8_t>(); if (!push_str(n)) return 2; break; }            // SHORT_BINUNICODE
            case 'X':  { if (!need(4)) return 2; const uint32_t n = rd<uint32_t>(); if (!push_str(n)) return 2; break; }          // BINUNICODE
            case 0x8d: { if (!need(8)) return 2; const uint64_t n = rd<uint64_t>(); if (!push_str(n)) return 2; break; }          // BINUNICODE8
            case 'C':  { if (!need(1)) return 2; const uint8_t n = rd<uint8_t>(); if (!skip_push_other(n)) return 2; break; }     // SHORT_BINBYTES
            case 'B':  { if (!need(4)) return 2; const uint32_t n = rd<uint32_t>(); if (!skip_push_other(n)) return 2; break; }   // BINBYTES
            case 0x8e: case 0x96: { if (!need(8)) return 2; const uint64_t n = rd<uint64_t>(); if (!skip_push_other(n)) return 2; break; }   // BINBYTES8 BYTEARRAY8
            case ']': case ')': stack.push_back(new_seq()); break;                   // EMPTY_LIST EMPTY_TUPLE
            case '}': case 0x8f: stack.push_back(IT_OTHER); break;                        // EMPTY_DICT EMPTY_SET
            case 'a': {                                                                       // APPEND
                if (stack.size() < 2) return 2;
                const Item x = stack.back(); stack.pop_back();
                if (is_seq(stack.back())) seqs[seq_index(stack.back())].push_back(x);
                break;
            }
            case 'e': {                                                                       // APPENDS
                const int64_t m = take_mark(); if (m < 1) return 2;
                const Item tgt = stack[(size_t)m - 1];
                if (is_seq(tgt)) {
                    std::vector<Item>& d = seqs[seq_index(tgt)];
                    d.insert(d.end(), stack.begin() + m, stack.end());
                }
                stack.resize((size_t)m);
                break;
            }
            case 'l': case 't': {                                                             // LIST TUPLE
                const int64_t m = take_mark(); if (m < 0) return 2;
                if (!seq_from_top(stack.size() - (size_t)m)) return 2;
                break;
            }
            case 0x85: if (!seq_from_top(1)) return 2; break;                                      // TUPLE1
            case 0x86: if (!seq_from_top(2)) return 2; break;                                      // TUPLE2
            case 0x87: if (!seq_from_top(3)) return 2; break;                                      // TUPLE3
            case 'd': case 0x91: {                                                            // DICT FROZENSET
                const int64_t m = take_mark(); if (m < 0) return 2;
                stack.resize((size_t)m); stack.push_back(IT_OTHER);
                break;
            }
            case 's': if (!pop(2)) return 2; if (stack.empty()) return 2; break;              // SETITEM
            case 'u': case 0x90: {                                                            // SETITEMS ADDITEMS
                const int64_t m = take_mark(); if (m < 1) return 2;
                stack.resize((size_t)m);
                break;
            }
            case 'c': {                                                                       // GLOBAL: two text lines
                for (int k = 0; k < 2; ++k) {
                    const void* nl = memchr(p, '\n', (size_t)(end - p));
                    if (!nl) return 2;
                    p = static_cast<const uint8_t*>(nl) + 1;
                }
                stack.push_back(IT_OTHER);
                break;
            }
            case 0x93: if (!pop(2)) return 2; stack.push_back(IT_OTHER); break;           // STACK_GLOBAL
            case 'R': case 0x81: if (!pop(2)) return 2; stack.push_back(IT_OTHER); break; // REDUCE NEWOBJ
            case 0x92: if (!pop(3)) return 2; stack.push_back(IT_OTHER); break;           // NEWOBJ_EX
            case 'b': {                                                                       // BUILD
                if (stack.size() < 2) return 2;
                const Item st = stack.back(); stack.pop_back();
                last_state = is_seq(st) ? seq_index(st) : -1;
                break;
            }
            case 'h': { if (!need(1)) return 2; const uint8_t i = rd<uint8_t>(); if (i >= memo.size()) return 2; stack.push_back(memo[i]); break; }     // BINGET
            case 'j': { if (!need(4)) return 2; const uint32_t i = rd<uint32_t>(); if (i >= memo.size()) return 2; stack.push_back(memo[i]); break; }   // LONG_BINGET
            case 'q': case 'r': {                                                             // BINPUT LONG_BINPUT
                uint32_t i;
                if (op == 'q') { if (!need(1)) return 2; i = rd<uint8_t>(); } else { if (!need(4)) return 2; i = rd<uint32_t>(); }
                if (stack.empty() || i > (1u << 30)) return 2;
                if (i >= memo.size()) memo.resize((size_t)i + 1, IT_OTHER);
                memo[i] = stack.back();
                break;
            }
            case 0x94: if (stack.empty()) return 2; memo.push_back(stack.back()); break;      // MEMOIZE
            default:
                return 1;
            }
        }
        return 2;                                                                             // ran off the end without STOP
    }
};

}  // namespace gm2tok
