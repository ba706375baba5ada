/* Smallest C client of libnerfds_b200.so: links against the C-ABI only (no CUDA headers, no C++),
 * checks the ABI version and struct sizes, and shows the error path of ndsr_create on a bad configuration.
 *   gcc -std=c99 -I include examples/abi_probe.c -L nerfds_b200/lib -lnerfds_b200 -Wl,-rpath,$PWD/nerfds_b200/lib -o abi_probe
 */
#include <stdio.h>
#include <string.h>

#include "nerfds_b200.h"

int main(void) {
  int32_t cfg_size = 0, ep_size = 0, out_size = 0;
  ndsr_struct_sizes(&cfg_size, &ep_size, &out_size);
  if (ndsr_abi_version() != NDSR_ABI_VERSION || cfg_size != (int32_t)sizeof(ndsr_config) ||
      ep_size != (int32_t)sizeof(ndsr_extra_params) || out_size != (int32_t)sizeof(ndsr_outputs)) {
    fprintf(stderr, "header / library mismatch\n");
    return 1;
  }
  ndsr_config cfg;
  memset(&cfg, 0, sizeof cfg);          /* size / abi_version left 0: must be rejected, not crash */
  ndsr_handle* h = NULL;
  const int rc = ndsr_create(&cfg, 0, &h);
  if (rc != NDSR_ERR_INVALID || h != NULL) {
    fprintf(stderr, "ndsr_create accepted an empty config (rc %d)\n", rc);
    return 2;
  }
  printf("abi %d ok; ndsr_create on an empty config: %d (%s)\n", ndsr_abi_version(), rc, ndsr_last_error(NULL));
  printf("sizeof config=%u extra_params=%u outputs=%u camera=%u tensor=%u ipc_handle=%u\n", (unsigned)sizeof(ndsr_config),
         (unsigned)sizeof(ndsr_extra_params), (unsigned)sizeof(ndsr_outputs), (unsigned)sizeof(ndsr_camera),
         (unsigned)sizeof(ndsr_tensor), (unsigned)sizeof(ndsr_ipc_handle));
  return 0;
}
