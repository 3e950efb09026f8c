#!/usr/bin/env python
"""oracle/patch_steps.py -- TEST INFRASTRUCTURE ONLY.

Writes a "patched reference" translation unit for configurations the stock reference cannot run
(BASELINE.md section 3): the number of diffusion sampling steps, hard-coded as the literals 80 / 79
and an 80-entry timestep table in diffusion() and diffusion_graph() (main.cpp:3084, 5641-5652, 5723,
5988-6032), becomes the run-time value HX_STEPS (environment variable of the same name, default 80);
the table becomes guided-diffusion's space_timesteps (accumulating form, SURVEY A-9), which
reproduces the 80 literals exactly.  Nothing else changes.  The output is a build intermediate
under oracle/_ref/ (git-ignored); oracle/Makefile deletes it after compiling.  Results produced
with it are labelled "patched reference", never as the stock reference.

usage: patch_steps.py <reference main.cpp> <out.cpp>
"""
import re
import sys

src = open(sys.argv[1]).read().split("\n")


def sub_range(lo, hi, pat, rep, expect):
    """regex substitution on 1-based inclusive line range; the match count is asserted"""
    n = 0
    for i in range(lo - 1, hi):
        src[i], k = re.subn(pat, rep, src[i])
        n += k
    assert n == expect, (pat, n, expect)


# diffusion_graph: number of time_embedding_<i> inputs
assert "int time_embedding_size = 80;" in src[3083]
src[3083] = src[3083].replace("= 80;", "= hx_steps();")
# diffusion(): the literal table -> computed table
assert src[5640].strip() == "std::vector<int> timestep_map = {" and src[5647].strip().endswith("3999};")
src[5640:5648] = ["  std::vector<int> timestep_map = hx_timestep_map();"] + [""] * 7
sub_range(5649, 5656, r"int diffusion_timesteps = 80;", "int diffusion_timesteps = hx_steps();", 1)
sub_range(5720, 5726, r"diffusion_index < 80;", "diffusion_index < hx_steps();", 1)
sub_range(5980, 6030, r"\b79 - diffusion_index", "(hx_steps() - 1) - diffusion_index", 8)
sub_range(6030, 6034, r"/ 80\.0\)", "/ (float)hx_steps())", 1)

prologue = r'''
// ---- patched-reference prologue (oracle/patch_steps.py) ----
#include <cmath>
#include <cstdlib>
#include <vector>
static int hx_steps() {
  static int n = -1;
  if (n < 0) { const char *e = getenv("HX_STEPS"); n = e ? atoi(e) : 80; }
  return n;
}
static std::vector<int> hx_timestep_map() {
  const int n = hx_steps();
  std::vector<int> m;
  const double frac = double(4000 - 1) / double(n - 1);
  double cur = 0.0;
  for (int i = 0; i < n; ++i) { m.push_back(int(std::round(cur))); cur += frac; }
  return m;
}
'''
open(sys.argv[2], "w").write(prologue + "\n".join(src))
