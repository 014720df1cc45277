#include <cstdio>
#include <cmath>
#include "vkgs_b200.h"
int main(int argc, char** argv)
{
  int ok = 0, bad = 0;
  for(int i = 1; i < argc; i++)
  {
    vkgs_scene* s = nullptr;
    int rc = vkgs_scene_load(argv[i], &s);
    if(rc == 0 && s)
    {
      vkgs_splat_set_view v;
      if(vkgs_scene_view(s, &v) == 0)
      {
        // touch everything the view exposes
        double acc = 0;
        for(uint64_t k = 0; k < v.count; k++)
        {
          acc += v.positions[3 * k] + v.positions[3 * k + 2] + v.f_dc[3 * k + 1] + v.opacity[k] + v.scale[3 * k + 2] + v.rotation[4 * k + 3];
          if(v.f_rest && v.f_rest_per_splat)
            acc += v.f_rest[k * v.f_rest_per_splat] + v.f_rest[(k + 1) * v.f_rest_per_splat - 1];
        }
        if(std::isnan(acc)) ok += 0;
      }
      vkgs_scene_free(s);
      ok++;
    }
    else
      bad++;
  }
  std::printf("loaded %d rejected %d\n", ok, bad);
  return 0;
}
