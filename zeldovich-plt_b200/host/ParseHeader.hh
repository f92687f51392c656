// ParseHeader-compatible parameter-file reader, written without flex/bison.
//
// Mirrors the public interface of the reference's vendored ParseHeader
// (reference subprojects/ParseHeader/include/ParseHeader.hh:14-90): HeaderStream,
// ParseHeader::installscalar / installvector / ReadHeader, MUST_DEFINE / DONT_CARE —
// for the grammar subset zeldovich parameter files use:
//   key = value [value ...] EOL          (phParser.yy:125-128)
//   '#' to end of line is a comment      (phScanner.ll:95)
//   a line starting with "##" toggles a block comment (phScanner.ll:98,163-173)
//   backslash-newline continues a statement (phScanner.ll:105)
//   include "file"                       (phScanner.ll:179-205)
//   the header ends at the byte pair 0x02 '\n' or at EOF (HeaderStream.cc:59-78)
//   unknown keys are ignored (ParseHeader.cc:30); a missing MUST_DEFINE key only warns
//   (phDriver.cc:369-379); int -> double promotes, double -> int truncates with a
//   warning (phDriver.cc:222-298).
// Numbers are scanned with the reference scanner's own token classes and converted with
// its own algorithm (phScanner.ll:136-145, myatod :274-301) so that every double equals
// what the reference would have parsed, bit for bit.
#pragma once

#include <cstdio>
#include <filesystem>
#include <map>
#include <string>
#include <variant>
#include <vector>

namespace fs = std::filesystem;

#define MUST_DEFINE true
#define DONT_CARE false

class HeaderStream {
public:
    explicit HeaderStream(const fs::path &fn);
    virtual ~HeaderStream();
    void OpenForRead();
    void Close();
    void ReadHeader();  // fills buffer with the header text, leaves fp at the end of the header

    fs::path name;
    char *buffer;
    size_t bufferlength;  // header length + 2 (the reference counts its terminator)
    FILE *fp;
};

void WriteHStream(FILE *fp, HeaderStream &in);

class ParseError : public std::runtime_error {
public:
    explicit ParseError(const std::string &m) : std::runtime_error(m) {}
};

class ParseHeader {
public:
    ParseHeader() {}
    virtual ~ParseHeader() {}

    template <typename T>
    void installscalar(const std::string &name, T &var, bool must_define) {
        Sym s;
        s.target      = &var;
        s.is_vector   = false;
        s.must_define = must_define;
        syms_[name]   = s;
        order_.push_back(name);
    }
    template <typename T>
    void installvector(const std::string &name, std::vector<T> &var, bool must_define, size_t maxlen = 1024) {
        Sym s;
        s.target      = &var;
        s.is_vector   = true;
        s.maxlen      = maxlen;
        s.must_define = must_define;
        syms_[name]   = s;
        order_.push_back(name);
    }

    // Parse the header of `in`; throws ParseError on a syntax or type error.
    void ReadHeader(HeaderStream &in);
    // Parse header text directly (used by ReadHeader and by the tests).
    void ParseText(const std::string &text, const std::string &origin);

    std::vector<std::string> warnings;  // "requires a value", truncation notes

private:
    typedef std::variant<int *, long long *, double *, std::string *, fs::path *, std::vector<int> *, std::vector<double> *>
       Target;
    struct Sym {
        Target target;
        bool is_vector   = false;
        bool must_define = false;
        bool seen        = false;
        size_t maxlen    = 1;
    };
    std::map<std::string, Sym> syms_;
    std::vector<std::string> order_;

public:
    // one scanned value
    struct Value {
        enum Kind { INT, FLOAT, STRING } kind;
        long long l;
        double d;
        std::string s;
    };

private:
    void assign(const std::string &key, Sym &sym, const std::vector<Value> &vals, const std::string &where);
    void parse_stream(const std::string &text, const std::string &origin, int depth);
};

// exposed for tests: the reference scanner's string -> double conversion
double ph_atod(const char *s);
