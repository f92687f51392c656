// zeldovich <param_file> — command-line front end with the reference's calling convention
// (reference src/zeldovich.cpp:848-852: exactly one argument, usage + exit(1) otherwise).
#include <cstdio>
#include <cstdlib>

#include "../../include/zeldovich_b200.h"

int main(int argc, char *argv[]) {
    if (argc != 2) {
        fprintf(stderr, "Usage: %s param_file\n", argv[0]);
        exit(1);
    }
    zplt_run_report rep;
    int rc = zplt_run_param_file(argv[1], -1, 1, &rep);
    if (rc != ZPLT_OK) {
        fprintf(stderr, "%s\n", zplt_last_error());
        exit(1);
    }
    return 0;
}
