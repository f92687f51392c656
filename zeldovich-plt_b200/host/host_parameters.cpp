#include "host_parameters.h"

#include <cmath>
#include <cstdio>

Parameters::Parameters(const fs::path &inputfile) : inputstream(nullptr) {
    // defaults (reference src/parameters.cpp:13-44)
    ppd = 0, numblock = 2, boxsize = 0, Pk_scale = 1;
    qdensity = 0, qascii = 0, qoneslab = -1;
    Pk_norm = 0, Pk_sigma = 0, Pk_sigma_ratio = 0, f_cluster = 1, Pk_smooth = 0;
    qPk_fix_to_mean = 0, seed = 0;
    Pk_filename = "", Pk_powerlaw_index = 1000;
    density_filename = "density{:d}";
    qonemode = 0, one_mode = {0, 0, 0};
    qPLT = 0, PLT_filename = "", qPLTrescale = 0, PLT_target_z = 0.;
    f_NL = 0., k_cutoff = 1., ICFormat = "", AllowDirectIO = 0, version = -1;
    CornerModes = 0, n_s = 1, Omega_M = 1.0;
    np = 0, cpd = 0, z_initial = 0;
    separation = fundamental = nyquist = 0;

    register_vars();
    inputstream = new HeaderStream(inputfile);
    ReadHeader(*inputstream);
    setup();
}

Parameters::~Parameters() {
    if (inputstream) {
        inputstream->Close();
        delete inputstream;
    }
}

// the 33 keys of reference src/parameters.cpp:61-95
void Parameters::register_vars() {
    installscalar("BoxSize", boxsize, MUST_DEFINE);
    installscalar("ZD_Pk_scale", Pk_scale, MUST_DEFINE);
    installscalar("NP", np, MUST_DEFINE);
    installscalar("ZD_NumBlock", numblock, MUST_DEFINE);
    installscalar("CPD", cpd, MUST_DEFINE);
    installscalar("ZD_qdensity", qdensity, DONT_CARE);
    installscalar("ZD_qoneslab", qoneslab, DONT_CARE);
    installscalar("ZD_Seed", seed, MUST_DEFINE);
    installscalar("ZD_Pk_norm", Pk_norm, MUST_DEFINE);
    installscalar("ZD_Pk_sigma", Pk_sigma, DONT_CARE);
    installscalar("ZD_Pk_sigma_ratio", Pk_sigma_ratio, DONT_CARE);
    installscalar("ZD_f_cluster", f_cluster, DONT_CARE);
    installscalar("ZD_Pk_smooth", Pk_smooth, MUST_DEFINE);
    installscalar("ZD_qPk_fix_to_mean", qPk_fix_to_mean, DONT_CARE);
    installscalar("ZD_Pk_filename", Pk_filename, DONT_CARE);
    installscalar("ZD_Pk_powerlaw_index", Pk_powerlaw_index, DONT_CARE);
    installscalar("InitialConditionsDirectory", output_dir, MUST_DEFINE);
    installscalar("ZD_density_filename", density_filename, DONT_CARE);
    installscalar("InitialRedshift", z_initial, MUST_DEFINE);
    installscalar("ZD_qonemode", qonemode, DONT_CARE);
    installvector("ZD_one_mode", one_mode, DONT_CARE);
    installscalar("ZD_qPLT", qPLT, DONT_CARE);
    installscalar("ZD_PLT_filename", PLT_filename, DONT_CARE);
    installscalar("ZD_qPLT_rescale", qPLTrescale, DONT_CARE);
    installscalar("ZD_PLT_target_z", PLT_target_z, DONT_CARE);
    installscalar("ZD_k_cutoff", k_cutoff, DONT_CARE);
    installscalar("ZD_f_NL", f_NL, DONT_CARE);
    installscalar("ZD_n_s", n_s, DONT_CARE);
    installscalar("Omega_M", Omega_M, DONT_CARE);
    installscalar("ICFormat", ICFormat, MUST_DEFINE);
    installscalar("AllowDirectIO", AllowDirectIO, DONT_CARE);
    installscalar("ZD_Version", version, DONT_CARE);
    installscalar("ZD_CornerModes", CornerModes, DONT_CARE);
}

#define REQUIRE(cond, msg) \
    do {                   \
        if (!(cond)) throw ParameterError(std::string("Invalid Parameters given: ") + (msg)); \
    } while (0)

// Checks and derived quantities of reference src/parameters.cpp:97-197, in its order.
int Parameters::setup() {
    if (version == -1)
        throw ParameterError(
           "*** ERROR: ZD_Version was not specified for zeldovich-PLT.  New ICs should specify ZD_Version = 2; legacy ICs "
           "(pre-November 2019) should use ZD_Version = 1 to reproduce the old phases.");
    REQUIRE(version == 1 || version == 2, "ZD_Version must be 1 or 2");
    REQUIRE(version == 2, "ZD_Version = 1 (GSL mt19937 phases that depend on ZD_NumBlock) is not supported by the B200 path");

    ppd = (int64_t) round(cbrt((double) np));
    fprintf(stderr, "Generating ICs for ppd = %lld\n", (long long) ppd);
    REQUIRE(ppd * ppd * ppd == np, "NP is not a perfect cube");
    REQUIRE(ppd <= MAX_PPD, "ppd exceeds 65536");

    REQUIRE(!(boxsize <= 0.0), "BoxSize must be positive");
    REQUIRE(!(ppd <= 0), "NP must be positive");
    REQUIRE(!(numblock <= 0), "ZD_NumBlock must be positive");
    REQUIRE(!(Pk_scale <= 0.0), "ZD_Pk_scale must be positive");
    REQUIRE(!(Pk_norm < 0.0), "ZD_Pk_norm must not be negative");
    if ((bool) (Pk_sigma > 0) == (bool) (Pk_sigma_ratio > 0)) throw ParameterError("Must specify exactly one of Pk_sigma or Pk_sigma_ratio!");
    REQUIRE(f_cluster > 0. && f_cluster <= 1., "ZD_f_cluster must lie in (0,1]");
    REQUIRE((!Pk_filename.empty()) != (bool) (Pk_powerlaw_index != 1000), "specify exactly one of ZD_Pk_filename and ZD_Pk_powerlaw_index");
    if (Pk_powerlaw_index != 1000) REQUIRE(Pk_powerlaw_index <= 0, "ZD_Pk_powerlaw_index must be <= 0");
    if (qPLT) REQUIRE(!PLT_filename.empty(), "ZD_qPLT needs ZD_PLT_filename");
    REQUIRE(k_cutoff >= 1, "ZD_k_cutoff must be >= 1");
    if (qPLT) REQUIRE(ICFormat.rfind("RV", 0) == 0, "ZD_qPLT needs an ICFormat that starts with RV");
    // BlockArray's constructor checks (reference src/block_array.cpp:38-40); ZD_NumBlock no
    // longer changes version-2 results but is still validated
    REQUIRE(ppd % 2 == 0, "PPD must be even");
    REQUIRE(numblock % 2 == 0, "ZD_NumBlock must be even");
    REQUIRE(ppd % numblock == 0, "ZD_NumBlock must divide PPD");

    separation  = boxsize / ppd;
    nyquist     = M_PI / separation;
    fundamental = 2.0 * M_PI / boxsize;

    if (qonemode) {
        REQUIRE(one_mode.size() >= 3, "ZD_one_mode needs three integers");
        fprintf(stderr, "one_mode: %d, %d, %d\n", one_mode[0], one_mode[1], one_mode[2]);
    }
    return 0;
}
