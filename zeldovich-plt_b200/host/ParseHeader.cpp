// Hand-written scanner + statement reader behind ParseHeader.hh.
#include "ParseHeader.hh"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

// ---------------------------------------------------------------- HeaderStream ----
HeaderStream::HeaderStream(const fs::path &fn) : name(fn), buffer(nullptr), bufferlength(0), fp(nullptr) {}
HeaderStream::~HeaderStream() {
    delete[] buffer;
    if (fp) fclose(fp);
}
void HeaderStream::OpenForRead() {
    if (name.empty()) throw ParseError("HeaderStream::OpenForRead: filename is empty");
    if (fp) throw ParseError("HeaderStream::OpenForRead: file is already open");
    fp = fopen(name.c_str(), "rb");
    if (!fp) throw ParseError("HeaderStream::OpenForRead:  cannot open filename \"" + name.string() + "\"");
}
void HeaderStream::Close() {
    if (fp) fclose(fp);
    fp = nullptr;
}
void HeaderStream::ReadHeader() {
    OpenForRead();
    std::string s;
    int c, prev = -1;
    bool terminated = false;
    while ((c = fgetc(fp)) != EOF) {
        if (prev == 0x2 && c == '\n') {
            s.pop_back();  // drop the 0x02
            terminated = true;
            break;
        }
        s.push_back((char) c);
        prev = c;
    }
    (void) terminated;
    delete[] buffer;
    bufferlength = s.size() + 2;
    buffer       = new char[bufferlength];
    memcpy(buffer, s.data(), s.size());
    buffer[bufferlength - 2] = 0;
    buffer[bufferlength - 1] = 0;
}
void WriteHStream(FILE *fp, HeaderStream &in) {
    if (in.buffer && in.bufferlength >= 2) fwrite(in.buffer, 1, in.bufferlength - 2, fp);
}

// ---------------------------------------------------------------- numbers ---------
// Same arithmetic as the reference scanner's myatod (phScanner.ll:274-301): digits are
// accumulated as val = 10*val + d across the decimal point, the fraction is undone by one
// division and the exponent applied with pow(10, e).  This is NOT strtod: e.g.
// "0.0210839935761" is 210839935761 / 1e13, which can differ from the correctly rounded
// value in the last bit, and parity with the reference needs the former.
double ph_atod(const char *s) {
    double val = 0.0, power = 1.0, eval = 0.0;
    int i = 0, sign, esign = 1;
    while (isspace((unsigned char) s[i])) i++;
    sign = (s[i] == '-') ? -1 : 1;
    if (s[i] == '-' || s[i] == '+') i++;
    for (; isdigit((unsigned char) s[i]); i++) val = 10.0 * val + (s[i] - '0');
    if (s[i] == '.') i++;
    for (; isdigit((unsigned char) s[i]); i++) {
        val = 10.0 * val + (s[i] - '0');
        power *= 10.0;
    }
    if (s[i] == 'e' || s[i] == 'E' || s[i] == 'd' || s[i] == 'D') {
        i++;
        esign = (s[i] == '-') ? -1 : 1;
        if (s[i] == '-' || s[i] == '+') i++;
        for (; isdigit((unsigned char) s[i]); i++) eval = 10.0 * eval + (s[i] - '0');
    }
    return (sign * val / power * pow(10.0, esign * eval));
}

namespace {

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_id_start(char c) { return isalpha((unsigned char) c) || c == '_' || c == '.' || c == '$'; }
inline bool is_id_char(char c) { return is_id_start(c) || is_digit(c); }

size_t digits(const char *s, size_t i) {
    size_t n = 0;
    while (is_digit(s[i + n])) n++;
    return n;
}

// length of the longest prefix of s matching the scanner's {int} class, 0 if none
size_t match_int(const char *s) {
    size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    size_t n = digits(s, i);
    return n ? i + n : 0;
}

// {exp1}: ((D|d|E|e)?(+|-) | (D|e|E)(+|-)?) digits     (phScanner.ll:139)
size_t match_exp1(const char *s) {
    size_t best = 0;
    {  // optional letter, mandatory sign
        size_t i = 0;
        if (s[i] == 'D' || s[i] == 'd' || s[i] == 'E' || s[i] == 'e') i++;
        if (s[i] == '+' || s[i] == '-') {
            size_t n = digits(s, i + 1);
            if (n) best = i + 1 + n;
        }
    }
    {  // mandatory letter (no lower-case d in the reference's class), optional sign
        if (s[0] == 'D' || s[0] == 'e' || s[0] == 'E') {
            size_t i = 1;
            if (s[i] == '+' || s[i] == '-') i++;
            size_t n = digits(s, i);
            if (n && i + n > best) best = i + n;
        }
    }
    return best;
}

// {float}: sign? mant1 exp1?  |  sign? mant2 exp2          (phScanner.ll:136-142)
size_t match_float(const char *s) {
    size_t i0   = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    size_t best = 0;
    {
        size_t a = digits(s, i0);
        if (s[i0 + a] == '.') {
            size_t b = digits(s, i0 + a + 1);
            if (a + b > 0) {
                size_t len = i0 + a + 1 + b;
                best       = len + match_exp1(s + len);
            }
        }
    }
    {
        size_t a = digits(s, i0);
        size_t i = i0 + a;
        if (a && (s[i] == 'D' || s[i] == 'd' || s[i] == 'E' || s[i] == 'e')) {
            i++;
            if (s[i] == '+' || s[i] == '-') i++;
            size_t n = digits(s, i);
            if (n && i + n > best) best = i + n;
        }
    }
    return best;
}

size_t match_id(const char *s) {
    if (!is_id_start(s[0])) return 0;
    size_t n = 1;
    while (is_id_char(s[n])) n++;
    return n;
}

struct Token {
    enum Kind { ID, VALUE, EQUALS, EOS, OTHER, INCLUDE } kind;
    ParseHeader::Value v;
    std::string text;
    int line;
};

}  // namespace

// ---------------------------------------------------------------- statements ------
void ParseHeader::assign(const std::string &key, Sym &sym, const std::vector<Value> &vals, const std::string &where) {
    auto type_error = [&](const char *t, const Value &v) {
        const char *vt = v.kind == Value::INT ? "INTEGER" : (v.kind == Value::FLOAT ? "DOUBLE" : "STRING");
        throw ParseError(where + ": attempt to set variable \"" + key + "\" of type " + t + " to value of type " + vt);
    };
    sym.seen = true;
    if (vals.size() > sym.maxlen)
        throw ParseError(where + ": number of values (" + std::to_string(vals.size()) + ") exceeds dimension for " + key + "[" +
                         std::to_string(sym.maxlen) + "]");
    if (sym.is_vector) {
        if (auto p = std::get_if<std::vector<int> *>(&sym.target)) {
            (*p)->clear();
            for (auto &v : vals) {
                if (v.kind == Value::STRING) type_error("INTEGER", v);
                if (v.kind == Value::FLOAT) warnings.push_back(where + ": truncating a float to an int for \"" + key + "\".");
                (*p)->push_back(v.kind == Value::INT ? (int) v.l : (int) v.d);
            }
        } else if (auto p = std::get_if<std::vector<double> *>(&sym.target)) {
            (*p)->clear();
            for (auto &v : vals) {
                if (v.kind == Value::STRING) type_error("DOUBLE", v);
                (*p)->push_back(v.kind == Value::INT ? (double) v.l : v.d);
            }
        }
        return;
    }
    const Value &v = vals[0];
    if (auto p = std::get_if<int *>(&sym.target)) {
        if (v.kind == Value::STRING) type_error("INTEGER", v);
        if (v.kind == Value::INT) {
            if (std::llabs(v.l) > std::numeric_limits<int>::max())
                throw ParseError(where + ": attempt to store too large a value: " + std::to_string(v.l) + " in an int variable: " + key);
            **p = (int) v.l;
        } else {
            warnings.push_back(where + ": truncating a float: " + std::to_string(v.d) + " to an int for \"" + key + "\".");
            **p = (int) v.d;
        }
    } else if (auto p = std::get_if<long long *>(&sym.target)) {
        if (v.kind == Value::STRING) type_error("LONG", v);
        if (v.kind == Value::FLOAT)
            warnings.push_back(where + ": truncating a float: " + std::to_string(v.d) + " to a long long int for \"" + key + "\".");
        **p = v.kind == Value::INT ? v.l : (long long) v.d;
    } else if (auto p = std::get_if<double *>(&sym.target)) {
        if (v.kind == Value::STRING) type_error("DOUBLE", v);
        **p = v.kind == Value::INT ? (double) v.l : v.d;
    } else if (auto p = std::get_if<std::string *>(&sym.target)) {
        if (v.kind != Value::STRING) type_error("STRING", v);
        **p = v.s;
    } else if (auto p = std::get_if<fs::path *>(&sym.target)) {
        if (v.kind != Value::STRING) type_error("PATH", v);
        **p = v.s;
    }
}

void ParseHeader::parse_stream(const std::string &text, const std::string &origin, int depth) {
    if (depth > 100) throw ParseError(origin + ": exceeded maximum include depth");
    const char *s = text.c_str();
    const size_t n = text.size();
    size_t i       = 0;
    int line       = 1;
    bool at_bol    = true;
    bool in_block  = false;
    std::vector<Token> stmt;

    auto where = [&](int ln) { return origin + ":" + std::to_string(ln); };

    auto flush = [&](int ln) {
        // one statement = tokens up to EOS
        if (stmt.empty()) return;
        std::vector<Token> t;
        t.swap(stmt);
        if (t[0].kind == Token::INCLUDE) {
            if (t.size() != 2 || t[1].kind != Token::VALUE || t[1].v.kind != Value::STRING)
                throw ParseError(where(ln) + ": include needs a quoted file name");
            FILE *f = fopen(t[1].v.s.c_str(), "r");
            if (!f) throw ParseError(where(ln) + ": failed to open include file \"" + t[1].v.s + "\". exiting...");
            std::string inc;
            char buf[4096];
            size_t got;
            while ((got = fread(buf, 1, sizeof(buf), f)) > 0) inc.append(buf, got);
            fclose(f);
            parse_stream(inc, t[1].v.s, depth + 1);
            return;
        }
        if (t[0].kind != Token::ID) throw ParseError(where(ln) + ": syntax error, unexpected value \"" + t[0].text + "\", expecting 'identifier ='");
        if (t.size() < 2 || t[1].kind != Token::EQUALS)
            throw ParseError(where(ln) + ": syntax error, unexpected string \"" + t[0].text + "\", expecting '='");
        if (t.size() < 3) throw ParseError(where(ln) + ": syntax error, no value after '=' for \"" + t[0].text + "\"");
        std::vector<Value> vals;
        for (size_t k = 2; k < t.size(); k++) {
            if (t[k].kind == Token::VALUE)
                vals.push_back(t[k].v);
            else if (t[k].kind == Token::ID) {
                Value v;
                v.kind = Value::STRING;
                v.s    = t[k].text;
                vals.push_back(v);
            } else
                throw ParseError(where(ln) + ": syntax error, unexpected \"" + t[k].text + "\" in value list of \"" + t[0].text + "\"");
        }
        auto it = syms_.find(t[0].text);
        if (it == syms_.end()) return;  // keys nobody registered are ignored
        assign(t[0].text, it->second, vals, where(ln));
    };

    while (i < n) {
        char c = s[i];
        if (at_bol && c == '#' && i + 1 < n && s[i + 1] == '#') {  // "##" at line start toggles a block comment
            in_block = !in_block;
            i += 2;
            at_bol = false;
            if (in_block) continue;
            // after leaving the block the rest of the line scans normally (it is usually empty)
            continue;
        }
        if (in_block) {
            if (c == '\n') {
                line++;
                at_bol = true;
            } else
                at_bol = false;
            i++;
            continue;
        }
        at_bol = false;
        if (c == '\n') {
            flush(line);
            line++;
            i++;
            at_bol = true;
            continue;
        }
        if (c == '#') {  // comment to end of line
            while (i < n && s[i] != '\n') i++;
            continue;
        }
        if (c == ' ' || c == '\t' || c == '\r') {
            i++;
            continue;
        }
        if (c == '\\') {  // continuation: backslash, optional blanks, newline
            size_t j = i + 1;
            while (j < n && (s[j] == ' ' || s[j] == '\t')) j++;
            if (j < n && s[j] == '\n') {
                line++;
                i = j + 1;
                continue;
            }
        }
        Token t;
        t.line = line;
        if (c == '"' || c == '\'') {
            size_t j = i + 1;
            while (j < n && s[j] != c && s[j] != '\n') j++;
            if (j >= n || s[j] != c) throw ParseError(where(line) + ": unterminated string");
            t.kind   = Token::VALUE;
            t.v.kind = Value::STRING;
            t.v.s    = text.substr(i + 1, j - i - 1);
            t.text   = t.v.s;
            i        = j + 1;
            stmt.push_back(t);
            continue;
        }
        // longest match among id / float / int, earlier class wins ties (flex rule order)
        size_t lid = match_id(s + i), lfl = match_float(s + i), lin = match_int(s + i);
        if (lid && lid >= lfl && lid >= lin) {
            t.text = text.substr(i, lid);
            t.kind = (t.text == "include" && stmt.empty()) ? Token::INCLUDE : Token::ID;
            i += lid;
        } else if (lfl && lfl >= lin) {
            t.text   = text.substr(i, lfl);
            t.kind   = Token::VALUE;
            t.v.kind = Value::FLOAT;
            t.v.d    = ph_atod(t.text.c_str());
            i += lfl;
        } else if (lin) {
            t.text   = text.substr(i, lin);
            t.kind   = Token::VALUE;
            t.v.kind = Value::INT;
            t.v.l    = atoll(t.text.c_str());
            i += lin;
        } else if (c == '=') {
            t.kind = Token::EQUALS;
            t.text = "=";
            i++;
        } else {
            t.kind = Token::OTHER;
            t.text = std::string(1, c);
            i++;
        }
        stmt.push_back(t);
    }
    flush(line);
}

void ParseHeader::ParseText(const std::string &text, const std::string &origin) {
    for (auto &kv : syms_) kv.second.seen = false;
    parse_stream(text, origin, 0);
    for (auto &name : order_) {
        Sym &s = syms_[name];
        if (s.must_define && !s.seen) {
            std::string w = "symbol \"" + name + "\" requires a value.";
            warnings.push_back(w);
            fprintf(stderr, "%s\n", w.c_str());
        }
    }
}

void ParseHeader::ReadHeader(HeaderStream &in) {
    in.ReadHeader();
    ParseText(std::string(in.buffer, in.bufferlength - 2), in.name.string());
}
