// Host-only check of the tensor-core engine's layer programs (no GPU needed): builds the op lists and assembles
// the pair programs of both levels for a config dumped by Python, verifies the ring schedule and prints the
// burst / step / MMA counts.   usage: tc_program_check cfg.bin
//   python -c "from nerfds_b200 import _lib; from nerfds_b200.config import nerf_ds_config; \
//              open('/tmp/cfg.bin','wb').write(bytes(_lib.to_c_config(nerf_ds_config(),'tc','split3')))"
#include "../nerfds_b200/csrc/nds_field_tc.cu"

#include <random>

using namespace nds;

static int pe_dim(int C, int lo, int hi, int ident) { return 2 * (hi - lo) * C + (ident ? C : 0); }

static void fill_dense(HostDense& d, int K, int N, std::mt19937& g) {
  std::uniform_real_distribution<float> u(-0.3f, 0.3f);
  d.K = K; d.N = N; d.W.resize((size_t)K * N); d.b.resize(N);
  for (auto& v : d.W) v = u(g);
  for (auto& v : d.b) v = u(g);
}
static void fill_mlp(HostMlp& m, int in_dim, int depth, int width, int skip, int out, std::mt19937& g) {
  m.depth = depth; m.width = width; m.in_dim = in_dim; m.skip = skip; m.hidden.resize(depth);
  int d = in_dim;
  for (int l = 0; l < depth; ++l) {
    if (l == skip) d += in_dim;
    fill_dense(m.hidden[l], d, width, g);
    d = width;
  }
  if (out > 0) fill_dense(m.logit, d, out, g);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s cfg.bin\n", argv[0]); return 2; }
  ndsr_handle* h = new ndsr_handle();
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(&h->cfg, sizeof(ndsr_config), 1, f) != 1) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  fclose(f);
  const ndsr_config& c = h->cfg;
  if (c.size != sizeof(ndsr_config)) { fprintf(stderr, "config size mismatch\n"); return 2; }
  h->H = c.use_hyper_sheet ? c.hyper_num_dims : 0;
  h->dim_mask_in = pe_dim(3, c.mask_min_deg, c.mask_max_deg, 0) + c.mask_embed_dims;
  h->dim_warp_in = pe_dim(3, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity) + c.warp_embed_dims + (c.use_mask_in_warp ? 1 : 0);
  h->dim_hyper_in = pe_dim(3, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg, 0) + c.warp_embed_dims + (c.use_mask_in_hyper ? 1 : 0);
  h->dim_trunk_in = pe_dim(3, c.spatial_min_deg, c.spatial_max_deg, c.use_posenc_identity) + (h->H ? pe_dim(h->H, c.hyper_point_min_deg, c.hyper_point_max_deg, 0) : 0);
  h->dim_view = c.use_viewdirs ? pe_dim(3, c.viewdir_min_deg, c.viewdir_max_deg, c.use_posenc_identity) : 0;
  h->dim_norm = c.norm_input_posenc ? pe_dim(3, c.norm_input_min_deg, c.norm_input_max_deg, c.use_posenc_identity) : 3;
  h->max_in = std::max(std::max(h->dim_trunk_in, h->dim_warp_in), std::max(h->dim_hyper_in, h->dim_mask_in));
  std::mt19937 g(1);
  HostModel& HM = h->host_model;
  fill_mlp(HM.mask, h->dim_mask_in, c.mask_depth, c.mask_width, c.mask_skip, 1, g);
  fill_mlp(HM.warp, h->dim_warp_in, c.warp_depth, c.warp_width, c.warp_skip, 0, g);
  fill_dense(HM.warp_w, c.warp_width, 3, g);
  fill_dense(HM.warp_v, c.warp_width, 3, g);
  fill_mlp(HM.hyper, h->dim_hyper_in, c.hyper_sheet_depth, c.hyper_sheet_width, c.hyper_sheet_skip, h->H, g);
  const int rgb_in = c.trunk_width + h->dim_view + (c.use_x_in_rgb_condition ? c.trunk_width : 0) + (c.predict_norm ? h->dim_norm : 0);
  for (int lv = 0; lv < 2; ++lv) {
    fill_mlp(HM.trunk[lv], h->dim_trunk_in, c.trunk_depth, c.trunk_width, c.trunk_skip, 0, g);
    fill_dense(HM.bottleneck[lv], c.trunk_width, c.trunk_width, g);
    fill_dense(HM.alpha[lv], c.trunk_width, c.predict_norm ? 4 : 1, g);
    fill_mlp(HM.rgb[lv], rgb_in, c.rgb_depth, c.rgb_width, -1, 3, g);
  }
  int rc = 0;
  for (int lv = 0; lv < 2; ++lv) {
    LevelBuild LB;
    Packed P;
    int r = build_level(h, lv, LB, P);
    if (r) { printf("level %d: build_level failed: %s\n", lv, h->err.c_str()); return 1; }
    static TcProgram prog;
    const char* names[4] = {"sigma-only", "full", "full carried", "full + grad"};
    for (int mode = 0; mode < 4; ++mode) {
      std::string err;
      if (!assemble(LB, mode != 0, mode == 2, mode == 3, 1024u, prog, err)) { printf("level %d %s: assemble failed: %s\n", lv, names[mode], err.c_str()); rc = 1; continue; }
      long mmas = 0, pair = 0;
      for (int i = 0; i < prog.n_burst; ++i) {
        const uint32_t ctl = prog.burst[i].ctl;
        const int steps = (ctl >> 15) & 7, pat = (ctl >> 13) & 3;
        const int per = pat == PAT_SS ? steps : 4;
        if (ctl & CTL_PAIR) { mmas += 2 * per; ++pair; }
        else mmas += per * ((ctl & B_TWO) ? 3 : 1);
      }
      printf("level %d %-13s ops %3d bursts %3d (pair %3ld) steps %3d mma/pair-of-tiles %ld  stream %.2f MB  smem %u B\n", lv, names[mode], prog.n_ops,
             prog.n_burst, pair, prog.n_steps, mmas, P.stream.size() / 1e6, prog.smem_bytes);
      if (getenv("DUMP") && mode == atoi(getenv("DUMP")) && lv == 1) {
        for (int i = 0; i < prog.n_burst; ++i) {
          const Burst& b = prog.burst[i];
          printf("  burst %3d slot %d rows %3d pat %d fl %04x pair %d dself %d unit %d d %3u a %u/%u\n", i, (b.ctl >> 18) & 1, ((b.idesc >> 17) & 63) * 8,
                 (b.ctl >> 13) & 3, b.ctl & 0x1fff, (b.ctl >> 25) & 1, (b.ctl >> 26) & 1, (b.ctl >> 19) & 3, b.d, b.a_hi, b.a_lo);
        }
        for (int i = 0; i < prog.n_steps; ++i) {
          const Step& st = prog.steps[i];
          const TcOp& op = prog.ops[st.op];
          printf("  step %3d kind %d slot %d op %3d arg %d | N %3d nnc %d epi %d dcol %3d mask %d/%d sg %d sd %d\n", i, st.kind, st.tslot, st.op, st.arg, op.N, op.n_nc, op.epi_kind,
                 op.d_col[0], op.mask_idx, op.mask_mode, op.signal_glue, op.signal_done);
        }
      }
    }
  }
  printf(rc ? "FAILED\n" : "ok\n");
  return rc;
}
