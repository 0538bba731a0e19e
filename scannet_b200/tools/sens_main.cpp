// `sens <file.sens> [outDir]` — drop-in for /root/reference/SensReader/c++/src/main.cpp (same stdout, same files).
#include "scannet_b200.h"
int main(int argc, const char** argv) { return scn_sens_main(argc, argv); }
