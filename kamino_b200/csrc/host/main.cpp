// main.cpp — `kamino-b200`: kamino's command line (src/main.rs:5-9) over the GPU front end.
#include "kamino_host.hpp"
int main(int argc, char** argv) { return kh::cli(argc, argv); }
