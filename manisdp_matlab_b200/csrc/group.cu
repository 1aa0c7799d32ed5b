// group.cu -- single-process multi-GPU handle: ONE caller thread (MATLAB's interpreter thread calls a MEX function
// synchronously, SURVEY 8b "Threading") drives G devices.
//
// The sharded layouts of the engine are SPMD: every rank issues the same sequence of C-ABI calls on its own handle and
// the calls meet in NCCL collectives, with host synchronisations in between (stop flags, RtrState read-backs).  One
// thread cannot play several ranks in turn -- it would block in rank 0's first read-back before rank 1 has enqueued its
// half of the collective -- so a group owns one WORKER THREAD per device.  Every worker creates a column-sharded
// ONLYUNITDIAG handle on its device (csrc/colshard.cu: all rows of C, ceil(p/G) columns of the factor while split; the
// NCCL id is made in-process) and then executes the closures the caller posts; a group call posts the same closure to
// all workers and returns when all are done.  The group keeps the handles MERGED between calls: set_Y / rand_Y / kkt /
// rank_cut / escape / line_search / get_Y run identically on every device (results are read from device 0), and
// manisdp_group_tr_solve is split -> tr_solve -> merge.
#include <condition_variable>
#include <functional>
#include <string.h>
#include <mutex>
#include <thread>

#include "common.cuh"

extern "C" int manisdp_nccl_unique_id(void* out128);

struct manisdp_group {
  int G = 0;
  std::vector<int> devices;
  std::vector<manisdp_t*> h;
  std::vector<std::thread> workers;
  std::vector<int> rc;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int, manisdp_t*&)> job;  // (rank, handle) -> status; null: exit
  uint64_t epoch = 0;
  int pending = 0;
  bool quit = false;
  std::string err;
};

static void group_worker(manisdp_group* g, int rank) {
  cudaSetDevice(g->devices[(size_t)rank]);
  uint64_t seen = 0;
  for (;;) {
    std::function<int(int, manisdp_t*&)> job;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_job.wait(lk, [&] { return g->quit || g->epoch != seen; });
      if (g->quit) return;
      seen = g->epoch;
      job = g->job;
    }
    const int rc = job(rank, g->h[(size_t)rank]);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->rc[(size_t)rank] = rc;
      if (--g->pending == 0) g->cv_done.notify_all();
    }
  }
}

// post `job` to every worker, wait for all; returns the first non-zero status (message of that rank kept)
static int group_run(manisdp_group* g, std::function<int(int, manisdp_t*&)> job) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->job = std::move(job);
    g->pending = g->G;
    g->epoch++;
  }
  g->cv_job.notify_all();
  {
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
  }
  for (int r = 0; r < g->G; ++r)
    if (g->rc[(size_t)r] != MANISDP_OK) {
      const char* m = manisdp_last_error(g->h[(size_t)r]);
      g->err = "device " + std::to_string(g->devices[(size_t)r]) + ": " + (m ? m : "");
      return g->rc[(size_t)r];
    }
  return MANISDP_OK;
}

extern "C" {

static std::string g_group_create_error;
const char* manisdp_group_last_error(const manisdp_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int manisdp_group_create(manisdp_group** out, const manisdp_problem* pb, int32_t ndev, const int32_t* devices) {
  if (!out || !pb || ndev < 1 || !devices) return MANISDP_E_ARG;
  *out = nullptr;
  if (pb->kind != MANISDP_ONLYUNITDIAG) return MANISDP_E_ARG;  // sharding exists for ONLYUNITDIAG
  manisdp_group* g = new manisdp_group();
  g->G = ndev;
  g->devices.assign(devices, devices + ndev);
  g->h.assign((size_t)ndev, nullptr);
  g->rc.assign((size_t)ndev, MANISDP_OK);
  unsigned char id[128];
  memset(id, 0, sizeof(id));
  if (ndev > 1 && manisdp_nccl_unique_id(id) != MANISDP_OK) {
    delete g;
    return MANISDP_E_NCCL;
  }
  for (int r = 0; r < ndev; ++r) g->workers.emplace_back(group_worker, g, r);
  const manisdp_problem base = *pb;
  const int rc = group_run(g, [g, base, &id](int rank, manisdp_t*& h) {
    manisdp_problem p = base;
    p.device = g->devices[(size_t)rank];
    p.rank = rank;
    p.world = g->G;
    p.row_begin = 0;
    p.row_end = p.n;
    p.nccl_unique_id = g->G > 1 ? (const void*)id : nullptr;
    p.shard_layout = MANISDP_SHARD_COLS;
    const int s = manisdp_create(&h, &p);
    return s;
  });
  if (rc != MANISDP_OK) {
    const char* m = manisdp_last_error(nullptr);  // (a failed manisdp_create leaves no handle to ask)
    g_group_create_error = std::string("group create: ") + (m ? m : "") + " " + g->err;
    // tear down what exists
    group_run(g, [](int, manisdp_t*& h) {
      if (h) manisdp_destroy(h);
      h = nullptr;
      return MANISDP_OK;
    });
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->quit = true;
    }
    g->cv_job.notify_all();
    for (auto& t : g->workers) t.join();
    delete g;
    return rc;
  }
  *out = g;
  return MANISDP_OK;
}

int manisdp_group_destroy(manisdp_group* g) {
  if (!g) return MANISDP_OK;
  group_run(g, [](int, manisdp_t*& h) {
    if (h) manisdp_destroy(h);
    h = nullptr;
    return MANISDP_OK;
  });
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->quit = true;
  }
  g->cv_job.notify_all();
  for (auto& t : g->workers) t.join();
  delete g;
  return MANISDP_OK;
}

int manisdp_group_size(const manisdp_group* g) { return g ? g->G : 0; }

int manisdp_group_set_Y(manisdp_group* g, const double* Y, int64_t p, int32_t layout) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int, manisdp_t*& h) { return manisdp_set_Y(h, Y, p, layout); });
}
int manisdp_group_rand_Y(manisdp_group* g, int64_t p, uint64_t seed) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int, manisdp_t*& h) { return manisdp_rand_Y(h, p, seed); });
}
int manisdp_group_get_Y(manisdp_group* g, double* Y, int32_t layout) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) { return rank == 0 ? manisdp_get_Y(h, Y, layout) : MANISDP_OK; });
}
int manisdp_group_get_stats(manisdp_group* g, manisdp_stats* out) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) { return rank == 0 ? manisdp_get_stats(h, out) : MANISDP_OK; });
}
int manisdp_group_cost(manisdp_group* g, double* f) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) {
    double v = 0.0;
    const int s = manisdp_cost(h, &v);
    if (rank == 0 && f) *f = v;
    return s;
  });
}
// trustregions(problem, Y, opts) on the column-split factor; info from device 0 (identical on all: the scalars of the
// loop are all-reduced).  hv_count counts products of the whole factor, as on one GPU.
int manisdp_group_tr_solve(manisdp_group* g, const manisdp_tr_options* opts, manisdp_tr_info* info) {
  if (!g) return MANISDP_E_ARG;
  manisdp_tr_options o;
  memset(&o, 0, sizeof(o));
  if (opts) o = *opts;
  return group_run(g, [=](int rank, manisdp_t*& h) {
    int s = manisdp_col_split(h);
    if (s != MANISDP_OK) return s;
    manisdp_tr_info mine;
    s = manisdp_tr_solve(h, &o, &mine);
    if (s != MANISDP_OK) return s;
    if (rank == 0 && info) *info = mine;
    return manisdp_col_merge(h);
  });
}
int manisdp_group_kkt(manisdp_group* g, int32_t delta, double eig_tol, int32_t update_dual, manisdp_kkt_info* out) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) {
    manisdp_kkt_info k;
    const int s = manisdp_kkt(h, delta, eig_tol, update_dual, &k);
    if (rank == 0 && out) *out = k;
    return s;
  });
}
int manisdp_group_rank_cut(manisdp_group* g, double theta, int32_t apply, int64_t* r, int64_t* p_new) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) {
    int64_t rr = 0, pp = 0;
    const int s = manisdp_rank_cut(h, theta, apply, &rr, &pp);
    if (rank == 0) {
      if (r) *r = rr;
      if (p_new) *p_new = pp;
    }
    return s;
  });
}
int manisdp_group_escape(manisdp_group* g, int32_t nne, double alpha, int32_t line_search) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int, manisdp_t*& h) { return manisdp_escape(h, nne, alpha, line_search); });
}
int manisdp_group_line_search(manisdp_group* g, double* alpha) {
  if (!g) return MANISDP_E_ARG;
  return group_run(g, [=](int rank, manisdp_t*& h) {
    double a = 0.0;
    const int s = manisdp_line_search(h, &a);
    if (rank == 0 && alpha) *alpha = a;
    return s;
  });
}

}  // extern "C"
