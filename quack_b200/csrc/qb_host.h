// qb_host.h -- internal host helpers shared by qb_api.cu and qb_host.cpp
#pragma once
#include <stdint.h>

#include <vector>

namespace qb {

// reference key order (first base most significant) -> kernel order (first base least significant)
uint32_t key_to_internal(uint32_t ref_key);

// exact 2^20-bit membership bitmap + exact 2^14-bit map of the 7-mer anchors (see qb_host.cpp) +
// linear-probing table of the keys (empty if more than a third of its slots would be used) for the
// kernels; n_anchors / anchor_density describe the anchor map (expected filter pass rate on random bases)
void build_adapter_images(const uint32_t *ref_keys, uint32_t n, std::vector<uint32_t> &bitmap,
                          std::vector<uint32_t> &anchor, std::vector<uint32_t> &exact, uint32_t &n_anchors,
                          double &anchor_density);

uint32_t gen_length(uint64_t seed, uint64_t i, uint32_t len_min, uint32_t len_max);
void gen_one(uint64_t seed, int mate, uint64_t i, uint32_t len, uint32_t len_max, double adapter_rate,
             uint8_t *seq, uint8_t *qual);

}  // namespace qb
