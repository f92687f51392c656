// Parameters: the control parameters of an IC run, read from a ParseHeader-style file.
// Same member names, defaults, registered keys and checks as the reference class
// (reference include/parameters.h:14-74, src/parameters.cpp:11-197); validation failures
// are reported by exception instead of assert/exit so that the C ABI can return them.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "ParseHeader.hh"

const int64_t MAX_PPD = 65536;  // reference include/zeldovich.h:34

class ParameterError : public std::runtime_error {
public:
    explicit ParameterError(const std::string &m) : std::runtime_error(m) {}
};

class Parameters : public ParseHeader {
public:
    double boxsize;
    double Pk_scale;
    int64_t ppd;
    int cpd;
    long long int np;
    int numblock;
    double separation, fundamental, nyquist;
    double k_cutoff;
    int qdensity, qascii, qoneslab;
    int seed;
    double Pk_norm, Pk_sigma, Pk_sigma_ratio;
    double f_cluster;
    double Pk_smooth;
    int qPk_fix_to_mean;
    fs::path Pk_filename;
    double Pk_powerlaw_index;
    fs::path output_dir;
    fs::path density_filename;
    double z_initial;
    HeaderStream *inputstream;
    int qonemode;
    std::vector<int> one_mode;
    int qPLT;
    fs::path PLT_filename;
    int qPLTrescale;
    double PLT_target_z;
    double f_NL, n_s, Omega_M;
    std::string ICFormat;
    int AllowDirectIO;
    int version;
    int CornerModes;

    explicit Parameters(const fs::path &inputfile);  // throws ParseError / ParameterError
    ~Parameters();
    void register_vars();
    int setup();  // derived quantities + checks; throws ParameterError
};
