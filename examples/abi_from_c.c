/* The boundary is a C ABI: this file is compiled as strict C99 against include/vkgs_b200.h (tests/test_abi.py) and
 * calls the host-only entry points; the compute entry points fail loudly without an sm_100 device. */
#include <stdio.h>
#include <stdlib.h>

#include "vkgs_b200.h"

int main(void)
{
  vkgs_options      opt;
  vkgs_camera       cam;
  vkgs_frame_params fp;
  vkgs_ctx*         ctx = NULL;
  int               rc;
  vkgs_default_options(&opt);
  vkgs_default_camera(&cam);
  rc = vkgs_frame_params_from_camera(&cam, 640u, 360u, &fp);
  printf("%s options=%u bytes frame_params=%u bytes focal=(%g, %g) rc=%d\n", vkgs_version(), vkgs_abi_struct_size(1),
         vkgs_abi_struct_size(2), fp.focal[0], fp.focal[1], rc);
  if(rc != VKGS_OK || vkgs_abi_struct_size(1) != (unsigned)sizeof(vkgs_options) || vkgs_abi_struct_size(2) != (unsigned)sizeof(vkgs_frame_params))
    return 1;
  rc = vkgs_create(0, &ctx);
  if(rc != VKGS_OK)
  {
    printf("vkgs_create: %d (no device: no CPU fallback)\n", rc);
    return rc == VKGS_ERR_NO_DEVICE ? 2 : 1;
  }
  vkgs_destroy(ctx);
  return 0;
}
