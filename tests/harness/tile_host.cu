// Host-side view of the packed-tile element order (csrc/common.cuh) for tests/test_tile_layout_cpu.py.
#include "../../magi_v2_b200/csrc/common.cuh"
extern "C" int tile_slot_host(int r, int cp) { return magi_tile_slot(r, cp); }
extern "C" int tile_pos_host(int r, int col) { return magi_tile_pos(r, col); }
