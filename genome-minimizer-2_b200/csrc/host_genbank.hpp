// host_genbank.hpp — part of libgm2.so (included by gm2.cu).  Host only.
//
// GenBank flat file -> the slice of the record the minimizer consumes:
//   sequence            `record.seq` after SeqIO.read(path, "genbank")          minimizer_2.py:455, :515, :35, :94
//   gene table          for every feature with type == "gene", in file order:
//                       name  = qualifiers.get("gene", [""])[0]                  :59-61
//                       start = int(location.start), end = int(location.end)     :78-79
// Biopython 1.85 (poetry.lock:4-5) is a third-party dependency of the reference and absent from its
// tree; the behaviour restated here is the one documented in SURVEY.md App. A and implemented by
// genome_minimizer_2_b200/genbank.py, which this scanner follows decision by decision (the tests
// compare the two, and the oracle's independent reader, on every fixture and on fuzzed files).
//
// Deliberately a SUBSET: anything unusual — zero or several records, carriage returns, non-ASCII or
// control bytes, remote / within-position / malformed locations, coordinates beyond 15 digits —
// is declined (Status::unsupported) and the caller uses the general Python reader, which also owns
// every error message.  Nothing here guesses.
#pragma once

#include <stdint.h>
#include <string>
#include <string_view>
#include <vector>

namespace gm2gb {

enum class Status { ok, unsupported };

struct Genes {
    std::string seq;                   // upper-case, blanks removed
    std::vector<int64_t> start, end;   // 0-based half-open span of each gene feature
    std::vector<int64_t> name_off;     // F + 1 offsets into `names`
    std::string names;                 // first /gene value of each gene feature ("" if none)
    int64_t n_features = 0;            // all feature-table entries, any key
    std::string why;                   // reason when declined
};

using sv = std::string_view;
static constexpr size_t npos = sv::npos;

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n'; }
inline sv lstrip(sv s) { size_t i = 0; while (i < s.size() && is_ws(s[i])) ++i; return s.substr(i); }
inline sv rstrip(sv s) { size_t n = s.size(); while (n > 0 && is_ws(s[n - 1])) --n; return s.substr(0, n); }
inline sv strip(sv s) { return rstrip(lstrip(s)); }
inline sv from(sv s, size_t i) { return i < s.size() ? s.substr(i) : sv(); }          // Python s[i:]
inline bool starts_with(sv s, sv p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(sv s, sv p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }

// ---- locations: strict grammar; whatever it accepts, genbank.parse_location maps to the same span
struct LocParser {
    sv t; size_t i = 0;
    int64_t lo = INT64_MAX, hi = INT64_MIN; int parts = 0;
    bool number(int64_t& v) {
        size_t j = i; v = 0;
        while (j < t.size() && t[j] >= '0' && t[j] <= '9') { v = v * 10 + (t[j] - '0'); ++j; }
        if (j == i || j - i > 15) return false;
        i = j; return true;
    }
    void fuzzy() { if (i < t.size() && (t[i] == '<' || t[i] == '>')) ++i; }
    bool simple() {
        int64_t a, b;
        fuzzy();
        if (!number(a)) return false;
        int64_t s = a - 1, e = a;                                         // "N"      -> [N-1, N)
        if (t.compare(i, 2, "..") == 0) {                                 // "N..M"   -> [N-1, M)
            i += 2; fuzzy();
            if (!number(b)) return false;
            e = b;
        } else if (i < t.size() && t[i] == '^') {                         // "N^M"    -> [N, N)
            ++i; fuzzy();
            if (!number(b)) return false;
            s = a; e = a;
        } else if (i < t.size() && t[i] == '.') return false;             // "N.M": the Python reader raises
        lo = s < lo ? s : lo; hi = e > hi ? e : hi; ++parts;
        return true;
    }
    bool loc(int depth) {
        if (depth > 8) return false;
        for (sv op : {sv("complement("), sv("join("), sv("order(")}) {
            if (t.compare(i, op.size(), op) == 0) {
                i += op.size();
                const bool list = op[0] != 'c';
                if (!loc(depth + 1)) return false;
                while (list && i < t.size() && t[i] == ',') { ++i; if (!loc(depth + 1)) return false; }
                if (i >= t.size() || t[i] != ')') return false;
                ++i; return true;
            }
        }
        return simple();
    }
};

inline bool parse_location(sv text, int64_t& start, int64_t& end) {
    std::string t;
    t.reserve(text.size());
    for (char c : text) if (!is_ws(c)) t.push_back(c);
    LocParser p; p.t = t;
    if (t.empty() || !p.loc(0) || p.i != t.size() || p.parts == 0) return false;
    start = p.lo; end = p.hi;
    return true;
}

// ---- one feature entry whose key is "gene": location span + first /gene value
inline sv qual_body(sv line) {                      // genbank._parse_feature_body: text of a continuation line
    return strip(line.substr(0, 21)).empty() ? from(line, 21) : strip(line);
}

inline bool parse_gene_entry(sv raw, int64_t& start, int64_t& end, std::string& name) {
    std::vector<sv> ch;                              // non-blank lines, key line first
    for (size_t a = 0; a <= raw.size();) {
        size_t b = raw.find('\n', a);
        if (b == npos) b = raw.size();
        const sv ln = raw.substr(a, b - a);
        if (!strip(ln).empty()) ch.push_back(ln);
        a = b + 1;
    }
    if (ch.empty()) return false;
    std::string loc(strip(from(ch[0], 21)));
    size_t i = 1;
    while (i < ch.size() && !starts_with(lstrip(from(ch[i], 21)), "/") && !starts_with(lstrip(ch[i]), "/")) {
        loc.append(strip(ch[i]));
        ++i;
    }
    if (!parse_location(loc, start, end)) return false;
    name.clear();
    while (i < ch.size()) {
        const sv body = qual_body(ch[i]);
        ++i;
        if (!starts_with(body, "/")) continue;                              // stray continuation
        const size_t eq = body.find('=');
        if (eq == npos) {                                                    // bare /key
            if (strip(body.substr(1)) == "gene") return true;                // first "gene" entry is "" (setdefault)
            continue;
        }
        const sv qk = body.substr(1, eq - 1);
        const sv qv0 = body.substr(eq + 1);
        std::string qv(qv0);
        if (starts_with(qv0, "\"")) {                                        // quoted, possibly over several lines
            sv last = qv0; size_t pieces = 1;
            while (((last == "\"" && pieces == 1) || !ends_with(last, "\"")) && i < ch.size()) {
                last = strip(qual_body(ch[i]));
                qv.push_back(' '); qv.append(last);
                ++i; ++pieces;
            }
        }
        if (qk != "gene") continue;
        sv v = qv;                                                           // genbank._unquote
        if (starts_with(v, "\"")) v.remove_prefix(1);
        if (ends_with(v, "\"")) v.remove_suffix(1);
        for (size_t k = 0; k < v.size(); ++k) {
            name.push_back(v[k]);
            if (v[k] == '"' && k + 1 < v.size() && v[k + 1] == '"') ++k;     // "" -> "
        }
        return true;
    }
    return true;                                                             // no /gene qualifier: ""
}

inline Status decline(Genes& g, const char* why) { g.why = why; return Status::unsupported; }

inline Status scan(sv text, Genes& g) {
    for (unsigned char c : text)
        if (c >= 0x7f || (c < 0x20 && c != '\n' && c != '\t')) return decline(g, "byte outside printable ASCII / tab / newline");
    // ---- exactly one LOCUS .. // record (genbank._split_records)
    size_t pos = starts_with(text, "LOCUS") ? 0 : text.find("\nLOCUS");
    if (pos == npos) return decline(g, "no record");
    if (text[pos] == '\n') ++pos;
    const size_t term = text.find("\n//", pos);
    size_t stop = text.size();
    if (term != npos) { const size_t nl = text.find('\n', term + 1); if (nl != npos) stop = nl + 1; }
    if (stop > 0 && text.find("\nLOCUS", stop - 1) != npos) return decline(g, "more than one record");
    const sv rec = text.substr(pos, stop - pos);

    // ---- feature table: from the line after FEATURES to the next line that starts in column 0
    const size_t f0 = starts_with(rec, "FEATURES") ? 0 : rec.find("\nFEATURES");
    if (f0 != npos) {
        const size_t nl = rec.find('\n', f0 + 1);
        if (nl != npos) {
            const size_t body0 = nl + 1;
            size_t stop_f = rec.size();
            for (size_t a = body0; a < rec.size();) {                      // first line start holding a non-blank
                if (!is_ws(rec[a])) { stop_f = a; break; }
                const size_t b = rec.find('\n', a);
                if (b == npos) break;
                a = b + 1;
            }
            const sv block = rec.substr(body0, stop_f - body0);
            // entries start at lines of the form: five blanks, then a non-blank (the key, columns 6-20)
            std::vector<size_t> starts;
            for (size_t a = 0; a < block.size();) {
                if (block.size() - a > 5 && block.compare(a, 5, "     ") == 0 && !is_ws(block[a + 5])) starts.push_back(a);
                const size_t b = block.find('\n', a);
                if (b == npos) break;
                a = b + 1;
            }
            starts.push_back(block.size());
            g.n_features = (int64_t)starts.size() - 1;
            g.name_off.push_back(0);
            std::string name;
            for (size_t k = 0; k + 1 < starts.size(); ++k) {
                const sv raw = block.substr(starts[k], starts[k + 1] - starts[k]);
                const sv keycols = raw.substr(5, 16);                        // raw[5:21]
                size_t e = 0;
                while (e < keycols.size() && !is_ws(keycols[e])) ++e;
                if (keycols.substr(0, e) != "gene") continue;
                int64_t s0, e0;
                if (!parse_gene_entry(raw, s0, e0, name)) return decline(g, "a gene feature's location is outside the supported grammar");
                g.start.push_back(s0); g.end.push_back(e0);
                g.names.append(name);
                g.name_off.push_back((int64_t)g.names.size());
            }
        }
    }
    if (g.name_off.empty()) g.name_off.push_back(0);

    // ---- sequence: ORIGIN lines from column 11, blanks removed, upper-cased
    const size_t o0 = rec.find("\nORIGIN");
    if (o0 != npos) {
        const size_t nl = rec.find('\n', o0 + 1);
        if (nl != npos) {
            const size_t body0 = nl + 1;
            size_t end = rec.find("\n//", body0 - 1);
            if (end == npos) end = rec.size();
            if (end > body0) {
                const sv block = rec.substr(body0, end - body0);
                g.seq.reserve(block.size());
                for (size_t a = 0; a <= block.size();) {
                    size_t b = block.find('\n', a);
                    if (b == npos) b = block.size();
                    for (size_t k = a + 10; k < b; ++k) {
                        const char c = block[k];
                        if (c != ' ') g.seq.push_back(c >= 'a' && c <= 'z' ? (char)(c - 32) : c);
                    }
                    a = b + 1;
                }
            }
        }
    }
    return Status::ok;
}

}  // namespace gm2gb
