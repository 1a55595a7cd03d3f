// The assembly plan object behind `femb_csr_plan*` (shared by assembly.cu and assembly_blocks.cu).
#pragma once
#include "common.cuh"

namespace femb {

// Block/colour plan of the fused P1 assembly (assembly_blocks.cu).  Nodes are clustered into blocks of R rows along a Morton
// curve of their coordinates; a block owns its rows' CSR values, lists every element touching one of its rows (sorted by
// colour: two elements of one colour share no OWNED node) and the nodes those elements reference beyond its own ("halo").
struct BlockPlan {
  int R = 0;                          // rows per block
  long long nblocks = 0, total = 0;   // blocks, sum of the blocks' element counts
  int max_colors = 0, max_halo = 0, max_acc = 0, max_elems = 0;
  int* blk_node = nullptr;            // [nblocks*R] global node of block row j (-1 = padding of the last block)
  int* row_gstart = nullptr;          // [nblocks*R] first CSR entry of the row (node_ptr[node])
  unsigned* row_meta = nullptr;       // [nblocks*R] offset in the block accumulator : 16 | row length : 8 | diagonal slot : 8
  int* blk_acc = nullptr;             // [nblocks] accumulator entries of the block
  int* blk_hptr = nullptr;            // [nblocks+1] halo lists
  int* blk_halo = nullptr;            // halo node ids, ascending per block
  int* blk_eptr = nullptr;            // [nblocks+1] element lists (colour-sorted, ascending element id within a colour)
  unsigned short* blk_cptr = nullptr; // [nblocks*(MAXC+1)] colour offsets relative to blk_eptr[b]
  uint4* rec = nullptr;               // [total] {local nodes 0|1<<16, 2|3<<16, slots of rows 0 and 1 (3 + 3 bytes, 2 spare)}
  unsigned* rec2 = nullptr;           // [total] slots of rows 2 and 3 ... see assembly_blocks.cu
  double build_ms = 0.0;
};
constexpr int BLK_MAXC = 64;
void block_plan_free(BlockPlan* b);

}  // namespace femb

struct femb_csr_plan {
  long long M = 0, N = 0, nnzn = 0;
  int nen = 0, max_row = 0, max_inc = 0;
  int* conn32 = nullptr;    // [M,nen]
  int* inc_ptr = nullptr;   // [N+1]
  int* inc = nullptr;       // [M*nen] flat slot e*nen+a, grouped by node, ascending
  int* node_ptr = nullptr;  // [N+1]
  int* node_col = nullptr;  // [nnzn] sorted within a row
  unsigned char* inc_slots = nullptr;  // [M*nen*nen] position of conn[e][b] in the row of the incidence's node (max_row <= 255)
  // P1 fused-assembly acceleration structure (built lazily): per tile of 32 consecutive rows, step-major records
  // rec[tile_ptr[t]*32 + step*32 + lane] = {other node 1, 2, 3, their three row slots packed in bytes}; x = -1 pads
  int4* rec = nullptr;
  int* tile_ptr = nullptr;          // [ntiles+1] in steps
  unsigned char* pdiag = nullptr;   // [N] slot of the diagonal entry
  long long ntiles = 0, total_steps = 0;
  femb::BlockPlan* blk = nullptr;   // block/colour plan of the fused P1 assembly (assembly_blocks.cu), built at the first call with coordinates
  bool blk_failed = false;          // a block exceeded the shared-memory budget: keep using the row-tile kernel
};

namespace femb {
int block_plan_build(femb_csr_plan* p, const double* coords, cudaStream_t s);
int block_assemble(femb_csr_plan* p, const double* coords, double* vals, int* flag, cudaStream_t s);
}  // namespace femb
