// Host-only report of the convolution planner (dynmm_b200/csrc/conv_plan.cuh) over every convolution of the
// ESANet-R34-NBt1D forward at 480x640: tiling, channel tile, rounds on the 148 SMs, pipeline depth, shared memory --
// and the tensor-pipe time the tiling implies against the ideal, i.e. how much of a launch is lost to tile
// quantisation before any latency is counted.  No GPU, no driver: tensor-map encoding is compiled out
// (-DDYNMM_PLAN_DRYRUN).  Build + run:  nvcc -std=c++17 -DDYNMM_PLAN_DRYRUN -I dynmm_b200/csrc -o tools/bin/plan_report
// tools/plan_report.cu && tools/bin/plan_report [batch] [active_depth_samples] [--csv]
//
// UMMA cost model (measured, tools/umma_issue_bench.cu): a 128 x N x 16 UMMA takes 64 cycles for N <= 128 and 128
// cycles for N = 256.  "ideal" = the same MMA work spread perfectly over all SMs at N = 128 efficiency.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "conv_plan.cuh"

namespace dynmm {
static char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int num_sms() { return 148; }
}  // namespace dynmm

using namespace dynmm;
using namespace dynmm::convk;

struct Layer {
  std::string name;
  int n, h, w, cin, cout, kh, kw, sh, sw;
  bool residual;
  int repeat;        // launches of this shape per forward (both encoders counted separately)
};

struct Row {
  Layer l;
  ConvPlan plan;
  int rc;
  long long tiles, rounds;
  double mma_cycles, ideal_cycles;
};

static Row plan_layer(const Layer& l, int sms, bool two_cta) {
  Row r{l, {}, 0, 0, 0, 0, 0};
  dynmm_conv_params p;
  memset(&p, 0, sizeof(p));
  static char fake[64] __attribute__((aligned(16)));
  p.in = p.weight = p.out = fake;
  p.residual = l.residual ? fake : nullptr;
  p.n = p.n_in = l.n;
  p.h_in = l.h;
  p.w_in = l.w;
  p.c_in = p.in_ld = l.cin;
  p.kh = l.kh;
  p.kw = l.kw;
  p.stride_h = l.sh;
  p.stride_w = l.sw;
  p.pad_h = l.kh / 2;
  p.pad_w = l.kw / 2;
  p.h_out = (l.h + 2 * p.pad_h - l.kh) / l.sh + 1;
  p.w_out = (l.w + 2 * p.pad_w - l.kw) / l.sw + 1;
  p.c_out = p.out_ld = p.res_ld = l.cout;
  p.relu = 1;
  r.rc = plan_conv(&p, &r.plan, sms, kSmemBudget, two_cta);
  if (r.rc) return r;
  const KernelArgs& a = r.plan.a;
  const int ctas = a.two_per_sm ? 2 * sms : sms;
  r.tiles = r.plan.max_tiles;
  r.rounds = ceil_div_ll(r.tiles, ctas);
  const int taps = l.kh * l.kw;
  const double umma_per_tile = (double)taps * a.k_chunks * (kBlockK / kUmmaK);
  const double cyc = a.tile_n > 128 ? 128.0 : 64.0;
  // two CTAs on one SM share its tensor pipe: a round of the pair costs two tiles
  r.mma_cycles = (double)r.rounds * umma_per_tile * cyc * (a.two_per_sm ? 2.0 : 1.0);
  const double pixels = (double)l.n * p.h_out * p.w_out;
  const int c_pad = (l.cout + 15) / 16 * 16;
  // ideal: every SM busy, full 128-row tiles, 64 cycles per 128 output channels (N <= 64 cannot beat 64 cycles either)
  const double n_units = c_pad <= 128 ? 1.0 : c_pad / 128.0;
  r.ideal_cycles = pixels / kBlockM * umma_per_tile * 64.0 * n_units / sms;
  return r;
}

int main(int argc, char** argv) {
  int batch = 8, active = -1;
  bool csv = false;
  std::vector<int> pos;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--csv")) csv = true;
    else pos.push_back(atoi(argv[i]));
  }
  if (pos.size() > 0) batch = pos[0];
  if (pos.size() > 1) active = pos[1];
  if (active < 0 || active > batch) active = batch;
  const int sms = 148;
  std::vector<Layer> layers;
  const int blocks[4] = {3, 4, 6, 3};
  const int planes[4] = {64, 128, 256, 512};
  for (int enc = 0; enc < 2; ++enc) {
    const int n = enc == 0 ? batch : active;
    if (n == 0) continue;
    const std::string e = enc == 0 ? "rgb" : "depth";
    int h = 120, w = 160, cin = 64;
    for (int s = 0; s < 4; ++s) {
      const int c = planes[s];
      const std::string st = e + " s" + std::to_string(s + 1);
      if (s > 0) {
        // first block: conv3x1 stride (2,1), conv1x3 stride (1,2), 1x1 stride-2 downsample, then 3x1 / 1x3 at the new size
        layers.push_back({st + " 3x1 s2 " + std::to_string(cin) + "->" + std::to_string(c), n, h, w, cin, c, 3, 1, 2, 1, false, 1});
        layers.push_back({st + " 1x3 s2", n, h / 2, w, c, c, 1, 3, 1, 2, false, 1});
        layers.push_back({st + " 1x1 s2 down", n, h, w, cin, c, 1, 1, 2, 2, false, 1});
        h /= 2;
        w /= 2;
        layers.push_back({st + " 3x1", n, h, w, c, c, 3, 1, 1, 1, false, 1 + 2 * (blocks[s] - 1)});
        layers.push_back({st + " 1x3 (+res)", n, h, w, c, c, 1, 3, 1, 1, true, 1 + 2 * (blocks[s] - 1)});
      } else {
        layers.push_back({st + " 3x1", n, h, w, c, c, 3, 1, 1, 1, false, 2 * blocks[s]});
        layers.push_back({st + " 1x3 (+res)", n, h, w, c, c, 1, 3, 1, 1, true, 2 * blocks[s]});
      }
      cin = c;
    }
  }
  // skip connections, context module, decoder (model.py:244-410, context_modules.py:47-87; nr_decoder_blocks 3,3,3)
  layers.push_back({"skip1 1x1 64->128", batch, 120, 160, 64, 128, 1, 1, 1, 1, false, 1});
  layers.push_back({"skip3 1x1 256->128", batch, 30, 40, 256, 128, 1, 1, 1, 1, false, 1});
  layers.push_back({"ppm final 1x1 768->128", batch, 15, 20, 768, 128, 1, 1, 1, 1, false, 1});
  const int dh[3] = {15, 30, 60}, dw[3] = {20, 40, 80};
  for (int i = 0; i < 3; ++i) {
    const std::string d = "dec" + std::to_string(i + 1);
    layers.push_back({d + " 3x3 c128", batch, dh[i], dw[i], 128, 128, 3, 3, 1, 1, false, 1});
    layers.push_back({d + " 3x1 c128", batch, dh[i], dw[i], 128, 128, 3, 1, 1, 1, false, 6});
    layers.push_back({d + " 1x3 c128 (+res)", batch, dh[i], dw[i], 128, 128, 1, 3, 1, 1, true, 6});
  }
  layers.push_back({"conv_out 3x3 128->40", batch, 120, 160, 128, 40, 3, 3, 1, 1, false, 1});

  if (csv) {
    printf("layer,launches,n,h,w,cin,cout,kh,kw,mode,b1,b2,bn,tile_n,c_tiles,tiles,ctas,rounds,stages,resident,two_per_sm,"
           "smem_kb,mma_cycles,ideal_cycles,efficiency\n");
  } else {
    printf("conv planner report: batch %d, %d active depth samples, %d SMs\n", batch, active, sms);
    printf("%-28s %3s %-5s %-9s %5s %6s %5s %3s %3s %4s %7s %7s %5s\n", "layer", "x", "mode", "box", "tileN", "tiles", "round",
           "stg", "res", "2cta", "mma_cyc", "ideal", "eff");
  }
  double tot_mma = 0, tot_ideal = 0;
  int launches = 0, failed = 0;
  for (const Layer& l : layers) {
    Row r = plan_layer(l, sms, true);
    if (r.rc) {
      printf("%-28s PLAN FAILED: %s\n", l.name.c_str(), g_err);
      ++failed;
      continue;
    }
    const KernelArgs& a = r.plan.a;
    const char* mode = a.tpg == 3 ? "halo" : (a.num_groups == 1 ? "1tap" : "taps");
    const double eff = r.ideal_cycles / r.mma_cycles;
    const int ctas = a.two_per_sm ? 2 * sms : sms;
    if (csv) {
      printf("%s,%d,%d,%d,%d,%d,%d,%d,%d,%s,%d,%d,%d,%d,%d,%lld,%d,%lld,%d,%d,%d,%.1f,%.0f,%.0f,%.3f\n", l.name.c_str(),
             l.repeat, l.n, l.h, l.w, l.cin, l.cout, l.kh, l.kw, mode, a.b1, a.b2, a.bn, a.tile_n, a.c_tiles, r.tiles, ctas,
             r.rounds, a.stages, a.b_resident, a.two_per_sm, r.plan.smem_bytes / 1024.0, r.mma_cycles, r.ideal_cycles, eff);
    } else {
      char box[32];
      snprintf(box, sizeof(box), "%dx%dx%d", a.b1, a.b2, a.bn);
      printf("%-28s %3d %-5s %-9s %5d %6lld %5lld %3d %3d %4d %7.0f %7.0f %5.2f\n", l.name.c_str(), l.repeat, mode, box,
             a.tile_n, r.tiles, r.rounds, a.stages, a.b_resident, a.two_per_sm, r.mma_cycles, r.ideal_cycles, eff);
    }
    tot_mma += r.mma_cycles * l.repeat;
    tot_ideal += r.ideal_cycles * l.repeat;
    launches += l.repeat;
  }
  if (!csv) {
    printf("TOTAL %d launches: tensor-pipe critical path %.0f cycles (%.1f us at 1.965 GHz), ideal %.0f cycles (%.1f us): "
           "tiling efficiency %.2f\n", launches, tot_mma, tot_mma / 1965.0, tot_ideal, tot_ideal / 1965.0, tot_ideal / tot_mma);
  }
  return failed ? 1 : 0;
}
