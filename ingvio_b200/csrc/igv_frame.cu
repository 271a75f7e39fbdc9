// igv_frame_step: one whole frame cycle of IngvioFilter::callbackMonoFrame / callbackStereoFrame
// (/root/reference/ingvio_estimator/src/IngvioFilter.cpp:143-231: propagateAugmentAtEnd -> visual update -> marginalise
// -> GNSS update) behind ONE C-ABI call.  In DEVICE pointer mode the kernel sequence of a frame is captured into a CUDA
// graph the second time the same frame shape is seen (same variable layout, ping-pong parity, argument pointers and
// sizes) and replayed from then on: a steady-state sliding window alternates between two graphs (the covariance
// ping-pong flips once per frame), the host submits one graph launch per frame instead of ~10 kernel launches and the
// bookkeeping of the variable layout is replayed from the snapshot taken at capture time.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "igv_internal.h"

namespace {

// Everything a frame's launches depend on besides device memory contents.
std::vector<unsigned long long> frame_key(const igv_batch* h, const igv_frame_args* a) {
  std::vector<unsigned long long> k;
  auto put = [&](unsigned long long v) { k.push_back(v); };
  auto putp = [&](const void* p) { k.push_back(reinterpret_cast<unsigned long long>(p)); };
  auto putd = [&](double d) { unsigned long long u; std::memcpy(&u, &d, 8); k.push_back(u); };
  put(h->cfg_version); put((unsigned long long)h->cur); put((unsigned long long)h->xcur); put((unsigned long long)h->N);
  for (const auto& v : h->vars) put(((unsigned long long)v.kind << 48) ^ ((unsigned long long)v.idx << 24) ^ ((unsigned long long)v.size << 8) ^ (unsigned long long)(v.tag & 0xff));
  for (int c : h->trk.col_of_slot) put(0x7000ull + (unsigned long long)c);
  put(0xfffffffffffffff0ull);
  put((unsigned long long)a->n_imu); putp(a->gyro); putp(a->accel); putp(a->dt);
  if (a->visual) {
    const igv_msckf_args* m = a->visual;
    put((unsigned long long)m->mode); put((unsigned long long)m->n_feats); put((unsigned long long)m->obs_slots);
    put((unsigned long long)(long long)m->max_valid); putd(m->noise);
    putp(m->pf_w); putp(m->anchor_slot); putp(m->obs); putp(m->obs_mask); putp(m->chi2_dof); putp(m->dx_out);
    putp(m->n_accepted_out); putp(m->gamma_out); putp(m->feat_ok);
  } else put(0ull);
  put((unsigned long long)a->n_marg);
  for (int i = 0; i < a->n_marg; ++i) put((unsigned long long)a->marg_slots[i]);
  if (a->gnss) {
    const igv_gnss_args* g = a->gnss;
    put((unsigned long long)g->n_sats); putp(g->unit); putp(g->res_pos); putp(g->res_vel); putp(g->sigma_psr);
    putp(g->sigma_dopp); putp(g->sys); putp(g->R_enu2ecef); putp(g->dx_out);
    put((unsigned long long)g->is_adjust_yof * 4 + g->chi2_test * 2 + g->strong_reject);
  } else put(0ull);
  return k;
}

igv_status run_frame(igv_batch* h, const igv_frame_args* a) {
  igv_status s;
  if (a->n_imu > 0) {
    if ((s = igv_propagate_imu(h, a->n_imu, a->gyro, a->accel, a->dt)) != IGV_OK) return s;
  }
  if ((s = igv_augment_clone(h)) != IGV_OK) return s;                       // propagateAugmentAtEnd (:143)
  if (a->visual && (s = igv_msckf_update(h, a->visual)) != IGV_OK) return s;  // :149-179
  for (int i = 0; i < a->n_marg; ++i)                                        // margSwPose (:175, :196)
    if ((s = igv_marginalize_clone(h, a->marg_slots[i])) != IGV_OK) return s;
  if (a->gnss && (s = igv_gnss_update(h, a->gnss)) != IGV_OK) return s;      // :201-229
  return IGV_OK;
}

void snapshot(const igv_batch* h, IgvFrameGraph& g, long long launches_before) {
  g.vars_after = h->vars; g.N_after = h->N; g.cur_after = h->cur; g.xcur_after = h->xcur;
  g.col_of_slot_after = h->trk.col_of_slot;
  g.launch_delta = h->launches - launches_before;
  g.last_visual_path = h->last_visual_path;
}

}  // namespace

extern "C" igv_status igv_frame_step(igv_batch* h, const igv_frame_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a || a->n_imu < 0 || a->n_marg < 0 || a->n_marg > IGV_MAX_CLONES) return IGV_ERR_INVALID;
  if (a->n_marg > 0 && !a->marg_slots) return IGV_ERR_INVALID;
  if (a->n_imu > 0 && (!a->gyro || !a->accel || !a->dt)) return IGV_ERR_INVALID;
  const bool graphable = h->ptr_mode == IGV_PTR_DEVICE && h->knobs.graph != 0 && !h->prof_on;
  if (!graphable) return run_frame(h, a);
  {   // errors left behind by earlier (unchecked) runtime calls must not be blamed on this frame
    cudaError_t stale = cudaGetLastError();
    if (stale != cudaSuccess && getenv("IGV_DEBUG")) std::fprintf(stderr, "igv_frame_step: stale error on entry: %s\n", cudaGetErrorString(stale));
  }
  const std::vector<unsigned long long> key = frame_key(h, a);
  IgvFrameGraph* hit = nullptr;
  for (auto& g : h->frame_graphs)
    if (g.key == key) { hit = &g; break; }
  if (hit && hit->exec) {   // replay: one graph launch + the host bookkeeping recorded at capture time
    cudaError_t e = cudaGraphLaunch(static_cast<cudaGraphExec_t>(hit->exec), h->stream);
    if (e != cudaSuccess) { h->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e); return IGV_ERR_CUDA; }
    h->vars = hit->vars_after; h->N = hit->N_after; h->cur = hit->cur_after; h->xcur = hit->xcur_after;
    h->trk.col_of_slot = hit->col_of_slot_after;
    h->launches += hit->launch_delta;
    h->last_visual_path = hit->last_visual_path;
    h->graph_replays++;
    hit->uses++;
    if (getenv("IGV_DEBUG")) {
      cudaError_t e2 = cudaStreamSynchronize(h->stream);
      std::fprintf(stderr, "igv_frame_step: replay N=%d cur=%d xcur=%d vars=%zu sync=%s last=%s\n", h->N, h->cur, h->xcur, h->vars.size(),
                   cudaGetErrorString(e2), cudaGetErrorString(cudaGetLastError()));
    }
    return IGV_OK;
  }
  if (!hit) {   // first sight of this frame shape: run it eagerly (workspaces grow, kernel attributes get set)
    if (h->frame_graphs.size() >= 8) {   // evict the least used entry
      size_t worst = 0;
      for (size_t i = 1; i < h->frame_graphs.size(); ++i) if (h->frame_graphs[i].uses < h->frame_graphs[worst].uses) worst = i;
      if (h->frame_graphs[worst].exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(h->frame_graphs[worst].exec));
      h->frame_graphs.erase(h->frame_graphs.begin() + worst);
    }
    IgvFrameGraph g;
    g.key = key;
    h->frame_graphs.push_back(g);
    return run_frame(h, a);
  }
  // second sight: capture, instantiate, launch
  const long long l0 = h->launches;
  igv_arena_quiesce(h);
  cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { cudaGetLastError(); return run_frame(h, a); }
  h->capturing = true;
  const igv_status s = run_frame(h, a);
  h->capturing = false;
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(h->stream, &graph);
  if (s != IGV_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (s != IGV_OK) return s;   // a failed call launched nothing during capture; the error stands
    h->err = "frame graph capture failed";
    return IGV_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return IGV_ERR_CUDA; }
  snapshot(h, *hit, l0);   // the host bookkeeping already advanced during capture
  hit->exec = exec;
  if (getenv("IGV_DEBUG")) std::fprintf(stderr, "igv_frame_step: captured N=%d cur=%d xcur=%d launches=%lld last=%s\n", h->N, h->cur, h->xcur,
                                        hit->launch_delta, cudaGetErrorString(cudaGetLastError()));
  e = cudaGraphLaunch(exec, h->stream);
  if (e != cudaSuccess) { h->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e); return IGV_ERR_CUDA; }
  h->graph_replays++;
  hit->uses++;
  return IGV_OK;
}

extern "C" long long igv_graph_replays(const igv_batch* h) { return h ? h->graph_replays : 0; }

void igv_frame_graphs_destroy(igv_batch* h) {
  for (auto& g : h->frame_graphs) if (g.exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.exec));
  h->frame_graphs.clear();
}
