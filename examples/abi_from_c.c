/* Calling libpea_b200.so from plain C: what a maintainer binding the library from C / cgo / JNI would write
 * (include/pea_b200.h is the whole contract: raw device pointers, sizes, a stream; no torch types).
 *
 *   gcc -std=c99 -Iinclude examples/abi_from_c.c -Lpea_diffusion_b200 -lpea_b200 -Wl,-rpath,$PWD/pea_diffusion_b200 -o abi_from_c
 *
 * Without a GPU it prints the library version and shows that a malformed call comes back as a negative PEA_E* code with a
 * message, not as a crash; with device pointers filled in (cudaMalloc or any framework's allocator) the same descriptor
 * launches the tcgen05 attention forward on the given stream.
 */
#include <stdio.h>
#include <string.h>

#include "pea_b200.h"

int main(void) {
  printf("pea_version %d (header %d)\n", pea_version(), PEA_B200_VERSION);
  if (pea_version() != PEA_B200_VERSION) return 1;

  pea_attn2_desc d;
  memset(&d, 0, sizeof d);
  d.B = 8; d.N = 1024; d.L = 77; d.H = 20; d.d = 64;   /* SDXL 32x32 block: train_sdxl_zh.py:397 */
  d.scale = 0.125f;
  d.dtype = PEA_BF16;
  d.ld_q = d.ld_o = 1280; d.ld_kv = 2560; d.bs_kv = 77 * 2560;
  d.flags = PEA_LAUNCH_PDL;
  /* d.d_q / d.d_k / d.d_v / d.d_o stay NULL here: the library must refuse, with text */
  int rc = pea_attn2_fwd(&d, NULL);
  printf("pea_attn2_fwd with null pointers -> %d: %s\n", rc, pea_last_error());
  if (rc >= 0) return 2;

  pea_gemm_desc g;
  memset(&g, 0, sizeof g);
  g.M = 1232; g.N = 1024; g.K = 1024;                   /* no operand planes / dtypes set: rejected before any launch */
  rc = pea_gemm(&g, NULL);
  printf("pea_gemm with a bad descriptor -> %d: %s\n", rc, pea_last_error());
  return rc < 0 ? 0 : 3;
}
