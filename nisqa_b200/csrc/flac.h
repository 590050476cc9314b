// flac.h - interface of the native FLAC reader (flac.cpp) used by the wav ingest (wavio.cpp).
#pragma once
#include <stdint.h>

#include <vector>

namespace nisqa {

struct FlacInfo {
  int32_t sample_rate = 0;
  int channels = 0, bits = 0, min_block = 0, max_block = 0;
  int64_t n_frames = 0;
};

bool flac_is(const char* path);                                   // file starts with the "fLaC" marker
bool flac_probe(const char* path, FlacInfo* info);                // STREAMINFO (length counted by decoding when it is not declared)
bool flac_decode_all(const char* path, FlacInfo* info, std::vector<int32_t>* interleaved);

}  // namespace nisqa
