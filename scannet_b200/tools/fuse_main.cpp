// `fuse <params.txt> [<params2.txt>] <file.sens> [out.ply]` — stands in for the reference's external
// DepthSensing.exe / FriedLiver.exe fusion stage (/root/reference/Server/scan_processor.py:27-35,123-138).
#include "scannet_b200.h"
int main(int argc, const char** argv) { return scn_fuse_main(argc, argv); }
