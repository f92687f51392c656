// TEST INFRASTRUCTURE — oracle build shim, not product code.
//
// Shadows the reference's subprojects/ParseHeader/include/ParseHeader.hh, whose
// detail/phDriver.hh needs a flex/bison-generated parser that cannot be produced
// in this image (no flex, no bison).  Provides the same public surface that
// reference src/parameters.cpp uses (HeaderStream, ParseHeader::installscalar /
// installvector / ReadHeader, WriteHStream, MUST_DEFINE / DONT_CARE) on top of a
// small hand-written reader for the `key = value...` subset that zeldovich
// parameter files use.  Number scanning follows the reference scanner's own
// rules (phScanner.ll:136-145 token classes, :274-301 myatod, atoll for ints) so
// that the doubles the reference would see are reproduced bit-for-bit.
#pragma once

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include <fmt/format.h>
#include <fmt/std.h>

namespace fs = std::filesystem;

#define MUST_DEFINE true
#define DONT_CARE false

class HeaderStream {
public:
    HeaderStream(const fs::path &fn) : name(fn), buffer(NULL), bufferlength(0), fp(NULL) {}
    virtual ~HeaderStream(void) { delete[] buffer; }

    void OpenForRead(void) {
        fp = fopen(name.c_str(), "rb");
        if (fp == NULL) {
            fmt::print(stderr, "HeaderStream::OpenForRead:  cannot open filename \"{}\"\n", name);
            exit(1);
        }
    }
    void Close(void) {
        if (fp != NULL) fclose(fp);
        fp = NULL;
    }
    // Header = everything up to the byte pair 0x02 '\n' or EOF
    // (reference HeaderStream.cc:59-78); +2 for the terminator the parser wants.
    void ReadHeader(void) {
        OpenForRead();
        std::string s;
        int c, prev = -1;
        while ((c = fgetc(fp)) != EOF) {
            if (prev == 0x2 && c == '\n') {
                s.pop_back();
                break;
            }
            s.push_back((char) c);
            prev = c;
        }
        bufferlength = s.size() + 2;
        buffer       = new char[bufferlength];
        memcpy(buffer, s.data(), s.size());
        buffer[bufferlength - 2] = 0;
        buffer[bufferlength - 1] = 0;
    }

    fs::path name;
    char *buffer;
    size_t bufferlength;
    FILE *fp;
};

inline void WriteHStream(FILE *fp, HeaderStream &in) {
    if (in.buffer) fwrite(in.buffer, 1, in.bufferlength - 2, fp);
}

class ParseHeader {
public:
    ParseHeader(void) {}
    ~ParseHeader() {}

    template <typename T>
    void installscalar(const std::string &name, T &var, bool must_define) {
        Sym s;
        s.ptr         = &var;
        s.is_vector   = false;
        s.must_define = must_define;
        s.seen        = false;
        syms[name]    = s;
    }

    template <typename T>
    void installvector(const std::string &name, std::vector<T> &var, bool must_define, size_t maxlen = 1024) {
        (void) maxlen;
        Sym s;
        s.ptr         = &var;
        s.is_vector   = true;
        s.must_define = must_define;
        s.seen        = false;
        syms[name]    = s;
    }

    void ReadHeader(HeaderStream &in) {
        in.ReadHeader();
        parse(std::string(in.buffer, in.bufferlength - 2));
        for (auto &kv : syms)
            if (kv.second.must_define && !kv.second.seen)
                fmt::print(stderr, "symbol \"{}\" requires a value.\n", kv.first);
    }

private:
    typedef std::variant<int *, long long *, double *, std::string *, fs::path *, std::vector<int> *, std::vector<double> *>
       VarPtr;
    struct Sym {
        VarPtr ptr;
        bool is_vector, must_define, seen;
    };
    std::unordered_map<std::string, Sym> syms;

    struct Val {
        enum { INT, DBL, STR } kind;
        long long l;
        double d;
        std::string s;
    };

    // reference phScanner.ll:274-301
    static double scan_double(const char *s) {
        double val, power, eval;
        int i = 0, sign, esign;
        sign = (s[i] == '-') ? -1 : 1;
        if (s[i] == '-' || s[i] == '+') i++;
        for (val = 0.0; isdigit((unsigned char) s[i]); i++) val = 10.0 * val + (s[i] - '0');
        if (s[i] == '.') i++;
        for (power = 1.0; isdigit((unsigned char) s[i]); i++) {
            val = 10.0 * val + (s[i] - '0');
            power *= 10.0;
        }
        if (s[i] == 'e' || s[i] == 'E' || s[i] == 'd' || s[i] == 'D') {
            i++;
            esign = (s[i] == '-') ? -1 : 1;
            if (s[i] == '-' || s[i] == '+') i++;
            for (eval = 0.0; isdigit((unsigned char) s[i]); i++) eval = 10.0 * eval + (s[i] - '0');
        } else {
            esign = 1;
            eval  = 0.0;
        }
        return (sign * val / power * pow(10.0, esign * eval));
    }

    static bool all_digits(const std::string &t, size_t from) {
        if (from >= t.size()) return false;
        for (size_t i = from; i < t.size(); i++)
            if (!isdigit((unsigned char) t[i])) return false;
        return true;
    }

    static Val classify(const std::string &t) {
        Val v;
        size_t p = (t[0] == '+' || t[0] == '-') ? 1 : 0;
        if (all_digits(t, p)) {
            v.kind = Val::INT;
            v.l    = atoll(t.c_str());
            return v;
        }
        // float: digits with '.', optional exponent; or digits with exponent
        bool numeric = p < t.size() && (isdigit((unsigned char) t[p]) || t[p] == '.');
        if (numeric) {
            bool ok = true, seen_digit = false;
            size_t i = p;
            while (i < t.size() && isdigit((unsigned char) t[i])) i++, seen_digit = true;
            if (i < t.size() && t[i] == '.') {
                i++;
                while (i < t.size() && isdigit((unsigned char) t[i])) i++, seen_digit = true;
            }
            if (!seen_digit) ok = false;
            if (ok && i < t.size()) {
                if (strchr("eEdD", t[i])) i++;
                if (i < t.size() && (t[i] == '+' || t[i] == '-')) i++;
                if (!all_digits(t, i)) ok = false;
            }
            if (ok) {
                v.kind = Val::DBL;
                v.d    = scan_double(t.c_str());
                return v;
            }
        }
        v.kind = Val::STR;
        v.s    = t;
        return v;
    }

    void assign(const std::string &key, Sym &sym, const std::vector<Val> &vals) {
        sym.seen = true;
        if (sym.is_vector) {
            if (auto pp = std::get_if<std::vector<int> *>(&sym.ptr)) {
                (*pp)->clear();
                for (auto &v : vals) (*pp)->push_back(v.kind == Val::INT ? (int) v.l : (int) v.d);
            } else if (auto pd = std::get_if<std::vector<double> *>(&sym.ptr)) {
                (*pd)->clear();
                for (auto &v : vals) (*pd)->push_back(v.kind == Val::INT ? (double) v.l : v.d);
            }
            return;
        }
        if (vals.empty()) return;
        const Val &v = vals[0];
        if (auto p = std::get_if<int *>(&sym.ptr)) {
            if (v.kind == Val::STR) goto bad;
            **p = v.kind == Val::INT ? (int) v.l : (int) v.d;
        } else if (auto p = std::get_if<long long *>(&sym.ptr)) {
            if (v.kind == Val::STR) goto bad;
            **p = v.kind == Val::INT ? (long long) v.l : (long long) v.d;
        } else if (auto p = std::get_if<double *>(&sym.ptr)) {
            if (v.kind == Val::STR) goto bad;
            **p = v.kind == Val::INT ? (double) v.l : v.d;
        } else if (auto p = std::get_if<std::string *>(&sym.ptr)) {
            if (v.kind != Val::STR) goto bad;
            **p = v.s;
        } else if (auto p = std::get_if<fs::path *>(&sym.ptr)) {
            if (v.kind != Val::STR) goto bad;
            **p = v.s;
        }
        return;
    bad:
        fmt::print(stderr, "ParseHeader shim: type mismatch for \"{}\"\n", key);
        exit(1);
    }

    void parse(const std::string &text) {
        // join continuation lines, strip comments, split statements at newlines
        std::vector<std::string> lines;
        std::string cur;
        bool in_block = false;
        size_t pos    = 0;
        while (pos <= text.size()) {
            size_t e         = text.find('\n', pos);
            std::string line = text.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            pos              = (e == std::string::npos) ? text.size() + 1 : e + 1;
            if (line.compare(0, 2, "##") == 0) {
                in_block = !in_block;
                continue;
            }
            if (in_block) continue;
            // strip # comments outside quotes
            char q = 0;
            for (size_t i = 0; i < line.size(); i++) {
                if (q) {
                    if (line[i] == q) q = 0;
                } else if (line[i] == '"' || line[i] == '\'')
                    q = line[i];
                else if (line[i] == '#') {
                    line.erase(i);
                    break;
                }
            }
            size_t last = line.find_last_not_of(" \t\r");
            if (last != std::string::npos && line[last] == '\\') {
                cur += line.substr(0, last) + " ";
                continue;
            }
            cur += line;
            lines.push_back(cur);
            cur.clear();
        }
        for (auto &ln : lines) {
            size_t eq = ln.find('=');
            if (eq == std::string::npos) continue;
            std::string key = ln.substr(0, eq);
            size_t a = key.find_first_not_of(" \t"), b = key.find_last_not_of(" \t\r");
            if (a == std::string::npos) continue;
            key     = key.substr(a, b - a + 1);
            auto it = syms.find(key);
            if (it == syms.end()) continue;  // unknown keys are ignored (ParseHeader.cc:30)
            std::vector<Val> vals;
            std::string rest = ln.substr(eq + 1);
            size_t i         = 0;
            while (i < rest.size()) {
                if (isspace((unsigned char) rest[i])) {
                    i++;
                    continue;
                }
                if (rest[i] == '"' || rest[i] == '\'') {
                    char q   = rest[i];
                    size_t j = rest.find(q, i + 1);
                    if (j == std::string::npos) j = rest.size();
                    Val v;
                    v.kind = Val::STR;
                    v.s    = rest.substr(i + 1, j - i - 1);
                    vals.push_back(v);
                    i = j + 1;
                } else {
                    size_t j = i;
                    while (j < rest.size() && !isspace((unsigned char) rest[j])) j++;
                    vals.push_back(classify(rest.substr(i, j - i)));
                    i = j;
                }
            }
            assign(key, it->second, vals);
        }
    }
};
