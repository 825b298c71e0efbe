// Descriptor codes of the band program, shared by the host builder (band_program.h) and the kernel (band_kernel.cu):
// code = op | variant << 4 | k << 6 | count << 9
#pragma once
namespace ppm {
enum {
  BD_TILE_BEGIN = 0, BD_TILE_END = 1, BD_INT = 2, BD_SGL = 3, BD_PAIR = 4, BD_END = 5,  // op
  BV_FULL = 0, BV_F = 1, BV_M = 2                                                       // variant
};
}  // namespace ppm
