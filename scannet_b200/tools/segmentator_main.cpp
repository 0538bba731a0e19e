// `segmentator input.ply [kThresh] [segMinVerts]` — drop-in for /root/reference/Segmentator/segmentator.cpp main().
#include "scannet_b200.h"
int main(int argc, const char** argv) { return scn_segmentator_main(argc, argv); }
