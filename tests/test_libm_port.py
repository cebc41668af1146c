"""CPU-only: csrc/ef_libm_f32.cuh (the sinf/cosf the HashSIFT kernel uses for the patch rotation) reproduces
the host libm bit for bit -- every 3rd float in [0, 2*pi*1.02] with both signs, plus a sample up to 3e38."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = r'''
#include "%s"
#include <cstdio>
#include <omp.h>
int main(){
  float hi = 6.4f; uint32_t hib; memcpy(&hib,&hi,4);
  long bad=0, cnt=0;
  #pragma omp parallel for reduction(+:bad,cnt) schedule(static)
  for (uint32_t b=0; b<=hib; b+=3){ for (int sg=0; sg<2; sg++){ float t; memcpy(&t,&b,4); if (sg) t=-t; cnt++;
    float c1=cosf(t), c2=ef_libm::cosf_glibc(t); if (memcmp(&c1,&c2,4)) bad++;
    float s1=sinf(t), s2=ef_libm::sinf_glibc(t); if (memcmp(&s1,&s2,4)) bad++; } }
  uint32_t b1; float big=3.0e38f; memcpy(&b1,&big,4);
  #pragma omp parallel for reduction(+:bad,cnt) schedule(static)
  for (uint32_t b=hib; b<=b1; b+=97){ float t; memcpy(&t,&b,4); cnt++;
    float c1=cosf(t), c2=ef_libm::cosf_glibc(t); if (memcmp(&c1,&c2,4)) bad++;
    float s1=sinf(-t), s2=ef_libm::sinf_glibc(-t); if (memcmp(&s1,&s2,4)) bad++; }
  printf("%%ld %%ld\n", cnt, bad); return 0; }
'''


def test_sinf_cosf_port_matches_host_libm(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC % (ROOT / "cuda-efficient-features_b200" / "csrc" / "ef_libm_f32.cuh"))
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", str(src), "-o", str(exe), "-lm"])
    cnt, bad = map(int, subprocess.check_output([str(exe)], timeout=600).split())
    assert cnt > 7e8 and bad == 0, (cnt, bad)
