// main.cpp — CPU ORACLE (test infrastructure): `kamino_oracle` CLI with the reference's flags
// (/root/reference/src/main.rs:5-9).
#include "kamino_oracle.hpp"
int main(int argc, char** argv) { return kmo::cli_main(argc, argv); }
