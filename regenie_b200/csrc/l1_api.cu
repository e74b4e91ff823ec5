// Level-1 ridge, LOCO assembly and Step-2 entry points of the C ABI (include/rg_b200.h).
#include <string.h>

#include <algorithm>

#include "context.cuh"

using namespace rg;

#define RG_API_BEGIN try {
#define RG_API_END                         \
  }                                        \
  catch (const rg::Error& e) {             \
    rg::set_last_error(e.msg);             \
    return 1;                              \
  }                                        \
  catch (const std::exception& e) {        \
    rg::set_last_error(e.what());          \
    return 1;                              \
  }                                        \
  return 0;



static void l1_fit(rg_ctx* h, const double* tau_host, double* cumsum, int32_t* best_idx) {
  RG_CHECK(h->kind == 1, "handle is not a Step-1 handle");
  RG_CHECK(h->R1 >= 1 && h->R1 <= kMaxRidge, "n_ridge_l1 out of range");
  RG_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  for (auto& l : h->lanes) RG_CUDA(cudaStreamSynchronize(l->stream));   // all level-0 blocks are in W
  const int K = h->K, R1 = h->R1, P = h->P;
  const int B = (int)h->B;
  const int loocv = h->loocv;
  const int nC = (int)round_up(B, 64), n_aug = nC + 64 + (loocv ? (int)h->Npad : 0), nmat = (loocv ? 1 : K) * R1;
  const int ldp = nC;
  const int64_t Npad = h->Npad;
  h->l1_nC = nC;
  // chunk table for the sample-axis reductions: bounded partial storage (<= ~1 GiB)
  {
    const int64_t per = (int64_t)nC * ldp * 8;
    const int64_t max_chunks = std::max<int64_t>(K, (1ll << 30) / per);
    int64_t len = round_up(std::max<int64_t>(kStatChunk, ceil_div(Npad, max_chunks - K + 1)), 128);
    std::vector<int4> chunks;
    std::vector<int2> fold_chunks(K);
    for (int f = 0; f < K; ++f) {
      fold_chunks[f].x = (int)chunks.size();
      for (int64_t o = 0; o < h->fold_pad_len[f]; o += len)
        chunks.push_back(make_int4((int)(h->fold_pad_start[f] + o),
                                   (int)std::min<int64_t>(len, h->fold_pad_len[f] - o), f, 0));
      fold_chunks[f].y = (int)chunks.size();
    }
    h->l1_nchunks = (int)chunks.size();
    h->l1_chunks.alloc(chunks.size());
    h->l1_fold_chunks.alloc(K);
    RG_CUDA(cudaMemcpyAsync(h->l1_chunks.p, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
    RG_CUDA(cudaMemcpyAsync(h->l1_fold_chunks.p, fold_chunks.data(), K * sizeof(int2), cudaMemcpyHostToDevice, s));
    RG_CUDA(cudaStreamSynchronize(s));
  }
  const int nch = h->l1_nchunks;
  const int64_t part_stride = (int64_t)nC * ldp;
  const int64_t cm_stride = (int64_t)n_aug * nC;
  const int ntiles = (int)(Npad / 128);
  const int NV = kMaxRidge * 3 + 2;
  h->l1_part.alloc((size_t)nch * part_stride);
  h->l1_part_y.alloc((size_t)nch * B);
  h->l1_cm.alloc((size_t)nmat * cm_stride);
  h->l1_inv.alloc(chol_inv_elems(nC, nmat));
  h->l1_beta.alloc((size_t)P * nmat * nC);
  h->l1_sums.alloc((size_t)P * NV);
  h->l1_part_out.alloc((size_t)ntiles * NV);
  h->l1_tau.alloc((size_t)P * R1);
  RG_CUDA(cudaMemcpyAsync(h->l1_tau.p, tau_host, (size_t)P * R1 * 8, cudaMemcpyHostToDevice, s));
  RG_CUDA(cudaMemsetAsync(h->l1_cm.p, 0, (size_t)nmat * cm_stride * 8, s));
  std::vector<int> loocv_best(P, 0);
  std::vector<double> loocv_cs((size_t)5 * P * R1, 0.0);
  if (loocv) {
    h->l1_zrows.alloc((size_t)P * Npad * nC);
    h->l1_hvec.alloc((size_t)P * Npad);
    h->l1_bvec.alloc((size_t)P * nC);
  }
  for (int p = 0; p < P; ++p) {
    const double* Wp = h->W.p + (size_t)p * Npad * h->B;
    const int ycol = h->C + p;
    launch_l1_gram(Wp, Npad, B, h->l1_chunks.p, nch, h->l1_part.p, part_stride, ldp, s);
    launch_l1_xty(Wp, Npad, h->xy.p, h->cpp, ycol, h->l1_chunks.p, nch, h->l1_part_y.p, B, s);
    launch_l1_assemble(h->l1_part.p, part_stride, ldp, h->l1_part_y.p, h->l1_fold_chunks.p, K, R1,
                       h->l1_tau.p + (size_t)p * R1, B, nC, h->l1_cm.p, cm_stride, loocv, s);
    if (loocv) launch_l1_loocv_fill(Wp, Npad, B, nC, h->l1_cm.p, cm_stride, nC + 64, R1, Npad, s);
    launch_chol_factor(h->l1_cm.p, cm_stride, nC, n_aug, nmat, h->l1_inv.p, h->err_slot.p, (long long)(1ll << 41) + p * 1024, s);
    if (loocv) {
      // CV sums of the closed-form LOO predictions for every tau, then the state rg_loco needs at tau*
      const int64_t inv_stride = (int64_t)(nC / 64) * 64 * 64;
      launch_l1_loocv_sums(h->l1_cm.p, cm_stride, nC, B, nC + 64, h->xy.p, h->cpp, ycol, h->l1_part_out.p, R1, ntiles,
                           h->l1_sums.p + (size_t)p * NV, s);
      std::vector<double> sv((size_t)R1 * 3);
      RG_CUDA(cudaMemcpyAsync(sv.data(), h->l1_sums.p + (size_t)p * NV, sv.size() * 8, cudaMemcpyDeviceToHost, s));
      double ne = 0.0;
      RG_CUDA(cudaMemcpyAsync(&ne, h->neff.p + p, 8, cudaMemcpyDeviceToHost, s));
      RG_CUDA(cudaStreamSynchronize(s));
      const double sy2 = ne - (double)h->C;                              // src/Step1_Models.cpp:891
      int bj = 0; double bv = 1e10;
      for (int j = 0; j < R1; ++j) {
        const double perf = (sv[3 * j + 1] + sy2 - 2 * sv[3 * j + 2]) / ne;
        if (perf < bv) { bv = perf; bj = j; }
      }
      loocv_best[p] = bj;
      for (int j = 0; j < R1; ++j) {
        loocv_cs[((size_t)0 * P + p) * R1 + j] = sv[3 * j];
        loocv_cs[((size_t)1 * P + p) * R1 + j] = 0.0;
        loocv_cs[((size_t)2 * P + p) * R1 + j] = sv[3 * j + 1];
        loocv_cs[((size_t)3 * P + p) * R1 + j] = sy2;
        loocv_cs[((size_t)4 * P + p) * R1 + j] = sv[3 * j + 2];
      }
      double* sys = h->l1_cm.p + (size_t)bj * cm_stride;
      double* rows = sys + (size_t)(nC + 64) * nC;
      launch_rows_sqnorm(rows, nC, B, h->l1_hvec.p + (size_t)p * Npad, ntiles, s);
      launch_chol_backsolve(sys, cm_stride, nC, 1, 1, h->l1_inv.p + (size_t)bj * inv_stride, s);   // b = H W^T y
      RG_CUDA(cudaMemcpyAsync(h->l1_bvec.p + (size_t)p * nC, sys + (size_t)nC * nC, (size_t)nC * 8,
                              cudaMemcpyDeviceToDevice, s));
      launch_chol_rows_backsolve(sys, cm_stride, nC, nC + 64, (int)Npad, 1, h->l1_inv.p + (size_t)bj * inv_stride, s);
      RG_CUDA(cudaMemcpyAsync(h->l1_zrows.p + (size_t)p * Npad * nC, rows, (size_t)Npad * nC * 8,
                              cudaMemcpyDeviceToDevice, s));
      h->launches += 6;
      continue;
    }
    launch_chol_backsolve(h->l1_cm.p, cm_stride, nC, 1, nmat, h->l1_inv.p, s);
    // keep beta[f][r][0:nC] (RHS row nC of every system)
    RG_CUDA(cudaMemcpy2DAsync(h->l1_beta.p + (size_t)p * nmat * nC, (size_t)nC * 8,
                              h->l1_cm.p + (size_t)nC * nC, (size_t)cm_stride * 8, (size_t)nC * 8, nmat,
                              cudaMemcpyDeviceToDevice, s));
    launch_l1_pred_sums(Wp, Npad, B, R1, h->l1_beta.p + (size_t)p * nmat * nC, nC, h->tile_fold.p, h->xy.p, h->cpp,
                        ycol, h->l1_part_out.p, ntiles, h->l1_sums.p + (size_t)p * NV, s);
    h->launches += 5 + chol_num_launches(nC) + 1 + 2;
  }
  if (loocv) {
    h->best_idx.assign(loocv_best.begin(), loocv_best.end());
    if (cumsum) memcpy(cumsum, loocv_cs.data(), loocv_cs.size() * 8);
    if (best_idx) for (int p = 0; p < P; ++p) best_idx[p] = loocv_best[p];
    h->l1_done = true;
    return;
  }
  std::vector<double> sums((size_t)P * NV), neff(P);
  RG_CUDA(cudaMemcpyAsync(sums.data(), h->l1_sums.p, sums.size() * 8, cudaMemcpyDeviceToHost, s));
  RG_CUDA(cudaMemcpyAsync(neff.data(), h->neff.p, P * 8, cudaMemcpyDeviceToHost, s));
  RG_CUDA(cudaStreamSynchronize(s));
  h->best_idx.assign(P, 0);
  for (int p = 0; p < P; ++p) {
    const double* v = &sums[(size_t)p * NV];
    double best = 1e10;                                   // src/Data.cpp:1021-1037
    for (int j = 0; j < R1; ++j) {
      const double sx = v[3 * j], sx2 = v[3 * j + 1], sxy = v[3 * j + 2], sy = v[3 * kMaxRidge], sy2 = v[3 * kMaxRidge + 1];
      if (cumsum) {
        cumsum[((size_t)0 * P + p) * R1 + j] = sx;
        cumsum[((size_t)1 * P + p) * R1 + j] = sy;
        cumsum[((size_t)2 * P + p) * R1 + j] = sx2;
        cumsum[((size_t)3 * P + p) * R1 + j] = sy2;
        cumsum[((size_t)4 * P + p) * R1 + j] = sxy;
      }
      const double perf = (sx2 + sy2 - 2 * sxy) / neff[p];
      if (perf < best) { best = perf; h->best_idx[p] = j; }
    }
    if (best_idx) best_idx[p] = h->best_idx[p];
  }
  h->l1_done = true;
}

static void loco(rg_ctx* h, const int32_t* chr_of_block, double* pred_out) {
  RG_CHECK(h->kind == 1 && h->l1_done, "rg_l1_fit must run before rg_loco");
  RG_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int K = h->K, R1 = h->R1, P = h->P, R = h->R;
  const int nC = h->l1_nC, nmat = K * R1;
  const int64_t Npad = h->Npad, N = h->N;
  // chromosomes in block order (blocks never straddle chromosomes, src/Data.cpp:311-334)
  std::vector<int32_t> chrs, col_start;
  for (int b = 0; b < h->total_blocks; ++b) {
    const int c = chr_of_block[b];
    RG_CHECK(c >= 1 && c <= 23, "chromosome out of range");
    if (chrs.empty() || chrs.back() != c) {
      RG_CHECK(chrs.empty() || c > chrs.back(), "blocks must be ordered by chromosome");
      chrs.push_back(c);
      col_start.push_back(b * R);
    }
  }
  col_start.push_back(h->total_blocks * R);
  const int nchr = (int)chrs.size();
  h->l1_chr_cols.alloc(col_start.size());
  RG_CUDA(cudaMemcpyAsync(h->l1_chr_cols.p, col_start.data(), col_start.size() * 4, cudaMemcpyHostToDevice, s));
  h->l1_pred.alloc((size_t)nchr * Npad);
  std::vector<double> pred((size_t)nchr * Npad);
  for (int p = 0; p < P; ++p) {
    if (h->loocv)
      launch_l1_loocv_chr_pred(h->W.p + (size_t)p * Npad * h->B, Npad, (int)h->B, nC, h->l1_zrows.p + (size_t)p * Npad * nC,
                               h->l1_hvec.p + (size_t)p * Npad, h->l1_bvec.p + (size_t)p * nC, h->xy.p, h->cpp,
                               h->C + p, nchr, h->l1_chr_cols.p, h->l1_pred.p, Npad, s);
    else
      launch_l1_chr_pred(h->W.p + (size_t)p * Npad * h->B, Npad, nchr, h->l1_chr_cols.p,
                         h->l1_beta.p + (size_t)p * nmat * nC, nC, R1, h->best_idx[p], h->tile_fold.p, h->l1_pred.p, Npad, s);
    h->launches += 1;
    RG_CUDA(cudaMemcpyAsync(pred.data(), h->l1_pred.p, pred.size() * 8, cudaMemcpyDeviceToHost, s));
    RG_CUDA(cudaStreamSynchronize(s));
    // LOCO assembly (src/Data.cpp:1846-1858): all-chromosome sum minus the chromosome's own part
    double* out = pred_out + (size_t)p * 23 * N;
    for (int64_t i = 0; i < N; ++i) {
      const int64_t t = h->pad_of[i];
      double tot = 0.0;
      for (int ci = 0; ci < nchr; ++ci) tot += pred[(size_t)ci * Npad + t];
      for (int c = 0; c < 23; ++c) out[(size_t)c * N + i] = tot;
      for (int ci = 0; ci < nchr; ++ci) out[(size_t)(chrs[ci] - 1) * N + i] = tot - pred[(size_t)ci * Npad + t];
    }
  }
}

extern "C" {

int rg_l1_fit(rg_handle h, const double* tau, double* cumsum, int32_t* best_idx) {
  RG_API_BEGIN
  RG_CHECK(h && tau, "null argument");
  l1_fit(h, tau, cumsum, best_idx);
  RG_CUDA(cudaGetLastError());
  RG_API_END
}

int rg_loco(rg_handle h, const int32_t* chr_of_block, double* pred_out) {
  RG_API_BEGIN
  RG_CHECK(h && chr_of_block && pred_out, "null argument");
  loco(h, chr_of_block, pred_out);
  RG_CUDA(cudaGetLastError());
  RG_API_END
}

int rg_W_info(rg_handle h, int32_t ph, void** dev_ptr, int64_t* ld, int64_t* ncols) {
  RG_API_BEGIN
  RG_CHECK(h && h->kind == 1 && ph >= 0 && ph < h->P, "bad argument");
  if (dev_ptr) *dev_ptr = h->W.p + (size_t)ph * h->Npad * h->B;
  if (ld) *ld = h->Npad;
  if (ncols) *ncols = h->B;
  RG_API_END
}

}  // extern "C"
