// Descriptor codes of the band program, shared by the host builder (band_program.h) and the kernel (band_kernel.cu):
// one 32-bit word per descriptor: op | variant << 4 | k << 6 | count << 9 (7 bits) | tile << 16 (tile: BD_TILE_BEGIN / BD_TILE_END)
#pragma once
namespace ppm {
enum {
  BD_TILE_BEGIN = 0, BD_TILE_END = 1,
  BD_REC = 2,    // 16-byte records (a, nbn): internal lineages (node array) and single leaves (entry bit 31), factor a + nbn Y'
  BD_PAIR = 4,   // 32-byte records: a static leaf pair as one quadratic in u
  BD_END = 5,
  BD_TILE_LP = 3,  // leaf-power program (band_program.h: leaf_pow): the tile's slices multiply in the product of ALL alive leaves as
                   // one power per slice; variant 1 (first, only with leaf heights != 0): height tables, variant 0: count tables
  BV_FULL = 0, BV_F = 1, BV_M = 2   // variant: every register / register k, all lanes / register k, lanes [lo, hi)
};
// stream layout: records per chunk -- 64 (two per lane; 32-byte leaf-pair records fit a ring slot), 128 for leaf-power programs (16-byte
// records only: half as many descriptors); a block = the chunk's entries + 4 words (the code first); the one-CTA layout uses 32.
// A record chunk's 7-bit count field holds count & 127: 0 stands for 128.
enum { kBandStreamChunk = 64, kBandStreamChunkLP = 128 };
static const unsigned kBandPadEntry = 0xffffffffu;  // stream layout: padding entry of a chunk (identity record)
}  // namespace ppm
