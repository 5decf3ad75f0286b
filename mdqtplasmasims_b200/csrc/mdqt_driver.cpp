// mdqt_driver.cpp -> mdqt_run: the time loop of the reference's main() (laserCoolingPlusExpansionMDQTSpeedUp.cpp:1139-1383)
// on top of the C ABI: same user inputs (by name, as options instead of edit-and-recompile), same directory tree,
// same output and restart files, hot path on the GPU.
//
//   mdqt_run <job>  [options]                     one trajectory, exactly `./runFile <job>` of the reference (SU:1145)
//   mdqt_run <job> --gpus G [options]             ONE large job (e.g. --N0 1000000) row-decomposed over G GPUs inside the library
//                                                 (mdqt_comm_init: NCCL all-gather of positions per MD step); same files
//   mdqt_run --jobs a-b [--batch 64] [--gpus 8]   the SLURM array of the reference (exampleSlurmFile.slurm:3,16:
//                                                 `--array=a-b`, `srun exe $SLURM_ARRAY_TASK_ID`) on one box: jobs a..b are
//                                                 dealt to the GPUs in contiguous blocks and advanced `batch` at a time in
//                                                 one handle; every job draws its own N in init() (SU:299-337), keeps its own
//                                                 seed (time + job, SU:1219, or --seed S -> S + job) and writes its own
//                                                 job directory, byte-compatible with a single run of that job.
//   options: [--Ge 0.1] [--density 2] [--sig0 4] [--Te 19] [--fracOfSig 0] [--N0 3500] [--detuning -1] [--detuningDP 1]
//            [--Om 1] [--OmDP 1] [--saveDirectory dataLaserCool/] [--newRun 1] [--c0 0] [--tmax 30]
//            [--reNormalizewvFns 0] [--sampleFreq 40] [--seed n] [--device 0] [--writers n] [--fast-single] [--quiet]
//
// A job gives the same bits whether it runs alone or inside any batch: the force summation order of a trajectory follows from
// its own ion count (mdqt_set_ion_counts with plan_n = 0), not from the batch, and the substep kernel's lane mappings are
// bitwise equivalent. (--fast-single is accepted and ignored: it used to lift a nominal-N0 plan for lone jobs.)
#include "../../include/mdqt.h"
#include "../../include/mdqt_io.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// output() formats ~9500 text lines per call and job (3 x 2001 velocity bins + one line per ion, SU:958-1024): about 3 ms
// of fprintf, as long as the 40 MD steps between two calls take on the GPU for ONE trajectory -- and 64 times that for a
// batch of 64. Writer threads do the formatting while the time loop keeps the GPU busy; a job's outputs always go to the
// same writer, in call order, so its files are byte-for-byte what a synchronous writer produces.
struct OutputJob {
  int kind = 0;  // 0: energies.dat line (appended: a job's lines stay in call order on one writer), 1: vel_dist files, 2: populations file
  std::string dir;
  unsigned counter = 0; int N = 0;
  mdqt_diag d; double Epot0 = 0;
  std::vector<double> pvel, pops, vx;  // vx = V[0][0..N): the only velocity row write_populations prints
};
class OutputWriter {
 public:
  OutputWriter() : th_([this] { run(); }) {}
  ~OutputWriter() { finish(); }
  void push(OutputJob&& j) {
    std::unique_lock<std::mutex> lk(m_);
    cv_space_.wait(lk, [this] { return q_.size() < 48; });  // bounded: at most 48 pieces in flight per writer
    q_.push_back(std::move(j));
    cv_work_.notify_one();
  }
  void finish() {
    { std::lock_guard<std::mutex> lk(m_); done_ = true; }
    cv_work_.notify_one();
    if (th_.joinable()) th_.join();
  }
 private:
  void run() {
    for (;;) {
      OutputJob j;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_work_.wait(lk, [this] { return done_ || !q_.empty(); });
        if (q_.empty()) return;
        j = std::move(q_.front());
        q_.pop_front();
        cv_space_.notify_one();
      }
      if (j.kind == 0) mdqt_io_append_energies(j.dir.c_str(), j.d.t, j.d.ekin_x, j.d.ekin_y, j.d.ekin_z, j.d.epot, j.Epot0, j.d.vx_avg);
      else if (j.kind == 1) mdqt_io_write_vel_dist(j.dir.c_str(), j.counter, j.pvel.data(), j.d.vx_avg);
      else mdqt_io_write_populations(j.dir.c_str(), j.counter, j.N, j.vx.data(), j.pops.data());
    }
  }
  std::mutex m_;
  std::condition_variable cv_work_, cv_space_;
  std::deque<OutputJob> q_;
  bool done_ = false;
  std::thread th_;
};
typedef std::vector<std::unique_ptr<OutputWriter>> Writers;
// one output() of one job = three independent pieces of text: the energies line goes to the job's own writer (order), the two bulky
// files (3 x 2001 + N lines) to the writers in turn, so that even a single job's formatting is spread over several threads
static std::atomic<unsigned> g_rr{0};
static void push_output(Writers& w, unsigned job, OutputJob&& full) {
  OutputJob a;
  a.kind = 0; a.dir = full.dir; a.d = full.d; a.Epot0 = full.Epot0;
  OutputJob b;
  b.kind = 1; b.dir = full.dir; b.counter = full.counter; b.d = full.d; b.pvel = std::move(full.pvel);
  OutputJob c;
  c.kind = 2; c.dir = std::move(full.dir); c.counter = full.counter; c.N = full.N; c.vx = std::move(full.vx); c.pops = std::move(full.pops);
  w[job % w.size()]->push(std::move(a));
  w[g_rr.fetch_add(1) % w.size()]->push(std::move(b));
  w[g_rr.fetch_add(1) % w.size()]->push(std::move(c));
}

struct Options {
  double Ge = 0.1, density = 2, sig0 = 4.0, Te = 19.0, fracOfSig = 0, detuning = -1, detuningDP = 1, Om = 1, OmDP = 1, tmax = 30;
  int N0 = 3500, newRun = 1, c0 = 0, renorm = 0, sampleFreq = 40, device = 0, writers = 0, fast_single = 0;
  bool quiet = false;
  long seed = 0;          // base seed
  bool seed_add_job = true;  // job j is seeded with seed + j (the reference's time(NULL) + job, SU:1219)
  std::string saveDirectory = "dataLaserCool/";
};

struct BatchResult { int ok = 0; long nsub = 0, nforce = 0, nout = 0; double wall = 0, ions = 0; };

static std::mutex g_print;
#define CKB(call)                                                                                        \
  do {                                                                                                   \
    if ((call) != 0) {                                                                                   \
      std::lock_guard<std::mutex> lk_(g_print);                                                          \
      fprintf(stderr, "mdqt_run: %s: %s\n", #call, mdqt_last_error());                                   \
      if (h) mdqt_destroy(h);                                                                            \
      return res;                                                                                        \
    }                                                                                                    \
  } while (0)

// jobs job0 .. job0+B-1 in ONE handle on `device`: init()/readConditions() per job, the main loop, output() per job,
// writeConditions() per job (SU:1139-1383)
static BatchResult run_batch(const Options& o, unsigned job0, int B, int device, std::vector<std::unique_ptr<OutputWriter>>& writers) {
  BatchResult res;
  mdqt_handle* h = NULL;
  const int ld = o.N0 + 1000;  // SU:126
  std::vector<std::string> dirs(B);
  std::vector<long> seeds(B);
  std::vector<int32_t> Nb(B);
  std::vector<unsigned> counter(B, 0);
  std::vector<double> R((size_t)B * 3 * ld), V((size_t)B * 3 * ld), psi_ld((size_t)ld * 24), tp_ld(ld);
  std::vector<std::vector<double>> psi_b(B), tp_b(B), vholder(B);
  double t = 0.0, L = 0, lDeb = 0;
  int c0 = o.c0;
  for (int b = 0; b < B; b++) {
    const unsigned job = job0 + b;
    char dir[1024];
    if (mdqt_io_dirname(dir, sizeof(dir), o.saveDirectory.c_str(), o.Ge, o.density, o.sig0, o.Te, o.fracOfSig, o.detuning,
                        o.detuningDP, o.Om, o.OmDP, o.N0, job, 1)) { fprintf(stderr, "mdqt_run: directory name too long\n"); return res; }
    dirs[b] = dir;
    seeds[b] = o.seed_add_job ? o.seed + (long)job : o.seed;
    vholder[b].assign((size_t)3 * MDQT_NUM_VINTERVALS * ld, 0.0);
    double* Rb = R.data() + (size_t)b * 3 * ld;
    double* Vb = V.data() + (size_t)b * 3 * ld;
    int N;
    if (o.newRun == 1) {
      N = mdqt_io_init_su(seeds[b], o.N0, o.Ge, ld, Rb, Vb, psi_ld.data(), tp_ld.data(), &L, &lDeb);  // init(), SU:289-348
      if (N < 0) { fprintf(stderr, "mdqt_run: job %u drew more than N0+1000 ions\n", job); return res; }
      { std::lock_guard<std::mutex> lk(g_print); printf("%i\n", N); }  // SU:338
    } else {
      double tb = 0;
      N = mdqt_io_read_conditions(dir, o.c0, ld, Rb, Vb, psi_ld.data(), &counter[b], &tb, vholder[b].data());  // SU:785-916
      if (N < 0) { fprintf(stderr, "mdqt_run: cannot read restart files for c0=%d in %s (%d)\n", o.c0, dir, N); return res; }
      std::fill(tp_ld.begin(), tp_ld.end(), 0.0);  // tPart is not checkpointed by the reference (Q7)
      t = tb;
    }
    Nb[b] = N;
    psi_b[b].assign(psi_ld.begin(), psi_ld.begin() + (size_t)N * 24);
    tp_b[b].assign(tp_ld.begin(), tp_ld.begin() + N);
  }
  if (o.newRun == 1) c0 = -1;  // SU:347
  const int Ncap = *std::max_element(Nb.begin(), Nb.end());
  // wavefunctions and tPart cross the ABI with the handle's capacity as their stride
  std::vector<double> psi((size_t)B * Ncap * 24, 0.0), tPart((size_t)B * Ncap, 0.0);
  for (int b = 0; b < B; b++) {
    std::copy(psi_b[b].begin(), psi_b[b].end(), psi.begin() + (size_t)b * Ncap * 24);
    std::copy(tp_b[b].begin(), tp_b[b].end(), tPart.begin() + (size_t)b * Ncap);
    for (int i = Nb[b]; i < Ncap; i++) psi[((size_t)b * Ncap + i) * 24] = 1.0;  // padding ions: a harmless normalised state
  }

  mdqt_params p;
  CKB(mdqt_params_su(&p, o.Ge, o.density, o.sig0, o.Te, o.fracOfSig, o.detuning, o.detuningDP, o.Om, o.OmDP, o.N0, Ncap));
  p.n_traj = B; p.traj0 = (int)job0; p.seed = (uint64_t)seeds[0]; p.device = device;
  p.renormalize = o.renorm;
  p.plan_n = 0;  // every trajectory's summation order follows from its own ion count: the same bits alone and in any batch
  CKB(mdqt_create(&p, &h));
  if (B > 1) {
    std::vector<uint64_t> s64(seeds.begin(), seeds.end());
    CKB(mdqt_set_ion_counts(h, Nb.data()));
    CKB(mdqt_set_traj_seeds(h, s64.data()));
  }
  CKB(mdqt_upload_state(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
  // RNG substep counter: continue the stream where a previous run of this job stopped
  const uint64_t sub0 = o.newRun == 1 ? 0 : (uint64_t)llround(t / p.dtq);
  CKB(mdqt_set_time(h, t, sub0));

  std::vector<double> Epot0(B, 0.0);
  CKB(mdqt_epot(h, Epot0.data()));  // Epotential(); Epot0 = Epot (SU:345-346). On resume the reference leaves Epot0 = 0 (Q8):
  if (o.newRun != 1) std::fill(Epot0.begin(), Epot0.end(), 0.0);

  std::vector<mdqt_diag> diag(B);
  std::vector<double> pvel((size_t)B * 3 * 2001), pops((size_t)B * Ncap * 3);
  int tsc = p.substeps_per_md;  // timeStepCounter (SU:1235)
  auto wall0 = std::chrono::steady_clock::now();
  for (;;) {
    int do_output, do_forces;
    int n = mdqt_schedule_next(&c0, &tsc, &t, p.substeps_per_md, o.sampleFreq, p.dtq, o.tmax, &do_output, &do_forces);
    if (n == 0) break;
    if (do_output) {  // output(), SU:917-1032
      CKB(mdqt_diagnostics(h, diag.data()));
      CKB(mdqt_vel_dist(h, pvel.data()));
      CKB(mdqt_populations(h, pops.data()));
      CKB(mdqt_download_state(h, NULL, V.data(), NULL, NULL, ld));
      for (int b = 0; b < B; b++) {
        OutputJob job;
        job.dir = dirs[b]; job.counter = counter[b]; job.N = Nb[b]; job.d = diag[b]; job.Epot0 = Epot0[b];
        job.pvel.assign(pvel.begin() + (size_t)b * 3 * 2001, pvel.begin() + (size_t)(b + 1) * 3 * 2001);
        job.pops.assign(pops.begin() + (size_t)b * Ncap * 3, pops.begin() + ((size_t)b * Ncap + Nb[b]) * 3);
        job.vx.assign(V.begin() + (size_t)b * 3 * ld, V.begin() + (size_t)b * 3 * ld + Nb[b]);
        push_output(writers, job0 + b, std::move(job));
        counter[b]++;
      }
      res.nout++;
    }
    if (do_forces && !do_output && n == p.substeps_per_md) {
      // whole MD steps with no output() in between: hand the run of them to mdqt_md_steps (one replayed CUDA graph)
      int k = 1;
      for (;;) {
        int c0b = c0, tscb = tsc, o2, f2;
        double tb = t;
        int n2 = mdqt_schedule_next(&c0b, &tscb, &tb, p.substeps_per_md, o.sampleFreq, p.dtq, o.tmax, &o2, &f2);
        if (n2 != p.substeps_per_md || !f2 || o2) break;
        c0 = c0b; tsc = tscb; t = tb; k++;
      }
      CKB(mdqt_md_steps(h, k));
      res.nforce += k; res.nsub += (long)k * n;
      continue;
    }
    if (do_forces) { CKB(mdqt_forces(h)); res.nforce++; }
    CKB(mdqt_substeps(h, n));
    res.nsub += n;
  }
  CKB(mdqt_download_state(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
  double t_dev; uint64_t s_dev;
  CKB(mdqt_get_time(h, &t_dev, &s_dev));
  if (t_dev != t) fprintf(stderr, "mdqt_run: warning: host/device clocks differ (%.17g vs %.17g)\n", t, t_dev);
  res.wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
  // every output file of these jobs must be on disk before their restart files are written: the caller drains the writers;
  // the restart files themselves do not depend on them
  for (int b = 0; b < B; b++) {
    if (mdqt_io_write_conditions(dirs[b].c_str(), c0, Nb[b], counter[b], R.data() + (size_t)b * 3 * ld, V.data() + (size_t)b * 3 * ld,
                                 psi.data() + (size_t)b * Ncap * 24, ld, vholder[b].data())) {  // SU:1381
      fprintf(stderr, "mdqt_run: cannot write restart files into %s\n", dirs[b].c_str());
      mdqt_destroy(h);
      return res;
    }
    res.ions += Nb[b];
  }
  if (!o.quiet) {
    std::lock_guard<std::mutex> lk(g_print);
    if (B == 1)
      fprintf(stderr, "mdqt_run: job %u, N=%d, t=%.6f, c0=%d: %ld substeps, %ld force calls, %ld outputs in %.3f s (%.3e ion-steps/s); files in %s\n",
              job0, Nb[0], t, c0, res.nsub, res.nforce, res.nout, res.wall, res.ions * res.nsub / res.wall, dirs[0].c_str());
    else
      fprintf(stderr, "mdqt_run: jobs %u-%u on GPU %d, N=%d..%d, t=%.6f, c0=%d: %ld substeps, %ld force calls, %ld outputs in %.3f s (%.3e ion-steps/s)\n",
              job0, job0 + B - 1, device, *std::min_element(Nb.begin(), Nb.end()), Ncap, t, c0, res.nsub, res.nforce, res.nout, res.wall,
              res.ions * res.nsub / res.wall);
  }
  mdqt_destroy(h);
  res.ok = 1;
  return res;
}

// ---- one LARGE job over G GPUs (BASELINE configs[4]): i-row decomposition inside the library (mdqt_comm_init), one host
// thread per GPU. Every thread drives the same schedule; the observables of output() are whole-system values on every
// rank (all-reduce inside the library), per-ion data are assembled in shared host arrays from each rank's own rows and
// written by rank 0 through the same writer as a single-GPU run: same files.
class ThreadBarrier {
 public:
  explicit ThreadBarrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const long gen = gen_;
    if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen_ != gen; });
  }
 private:
  std::mutex m_; std::condition_variable cv_; int n_, count_ = 0; long gen_ = 0;
};

static int run_rows(const Options& o, unsigned job, int G, std::vector<std::unique_ptr<OutputWriter>>& writers) {
  const int ld = o.N0 + 1000;  // SU:126
  char dir[1024];
  if (mdqt_io_dirname(dir, sizeof(dir), o.saveDirectory.c_str(), o.Ge, o.density, o.sig0, o.Te, o.fracOfSig, o.detuning, o.detuningDP,
                      o.Om, o.OmDP, o.N0, job, 1)) { fprintf(stderr, "mdqt_run: directory name too long\n"); return 1; }
  const long seed = o.seed_add_job ? o.seed + (long)job : o.seed;
  std::vector<double> R((size_t)3 * ld), V((size_t)3 * ld), psi((size_t)ld * 24), tPart(ld, 0.0), vholder((size_t)3 * MDQT_NUM_VINTERVALS * ld, 0.0);
  unsigned counter = 0;
  double t0 = 0.0, L = 0, lDeb = 0;
  int N, c0_start = o.c0;
  if (o.newRun == 1) {
    N = mdqt_io_init_su(seed, o.N0, o.Ge, ld, R.data(), V.data(), psi.data(), tPart.data(), &L, &lDeb);
    if (N < 0) { fprintf(stderr, "mdqt_run: more than N0+1000 ions drawn\n"); return 1; }
    printf("%i\n", N);
    c0_start = -1;
  } else {
    N = mdqt_io_read_conditions(dir, o.c0, ld, R.data(), V.data(), psi.data(), &counter, &t0, vholder.data());
    if (N < 0) { fprintf(stderr, "mdqt_run: cannot read restart files for c0=%d in %s (%d)\n", o.c0, dir, N); return 1; }
  }
  if (G > N) G = N;
  const int rows = (N + G - 1) / G;
  while ((G - 1) * rows >= N) G--;  // every rank owns at least one row
  unsigned char uid[128];
  if (mdqt_comm_unique_id(uid)) { fprintf(stderr, "mdqt_run: %s\n", mdqt_last_error()); return 1; }
  std::vector<double> pops((size_t)N * 3), pvel(3 * 2001);
  ThreadBarrier bar(G);
  std::vector<int> rc(G, 0);
  int c0_end = 0; double t_end = 0; long nsub_total = 0, nforce = 0, nout = 0; double wall = 0;
  auto worker = [&](int g) {
    mdqt_handle* h = NULL;
    auto bail = [&](const char* what) { fprintf(stderr, "mdqt_run: rank %d: %s: %s\n", g, what, mdqt_last_error()); rc[g] = 1; };
#define CKR(call) do { if ((call) != 0) { bail(#call); if (h) mdqt_destroy(h); return; } } while (0)
    mdqt_params p;
    CKR(mdqt_params_su(&p, o.Ge, o.density, o.sig0, o.Te, o.fracOfSig, o.detuning, o.detuningDP, o.Om, o.OmDP, o.N0, N));
    p.traj0 = (int)job; p.seed = (uint64_t)seed; p.device = o.device + g; p.renormalize = o.renorm;
    p.row0 = g * rows; p.n_rows = std::min(rows, N - g * rows);
    p.plan_n = 0;  // the same bits as the one-GPU run of this job
    CKR(mdqt_create(&p, &h));
    CKR(mdqt_comm_init(h, uid, g, G));
    CKR(mdqt_upload_state(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
    const uint64_t sub0 = o.newRun == 1 ? 0 : (uint64_t)llround(t0 / p.dtq);
    CKR(mdqt_set_time(h, t0, sub0));
    double Epot0 = 0.0;
    CKR(mdqt_epot(h, &Epot0));
    if (o.newRun != 1) Epot0 = 0.0;  // Q8
    int c0 = c0_start, tsc = p.substeps_per_md;
    double t = t0;
    long nsub = 0, nf = 0, no = 0;
    std::vector<double> pops_rows((size_t)p.n_rows * 3), pv(3 * 2001);
    auto wall0 = std::chrono::steady_clock::now();
    for (;;) {
      int do_output, do_forces;
      int n = mdqt_schedule_next(&c0, &tsc, &t, p.substeps_per_md, o.sampleFreq, p.dtq, o.tmax, &do_output, &do_forces);
      if (n == 0) break;
      if (do_output) {
        mdqt_diag d;
        CKR(mdqt_diagnostics(h, &d));
        CKR(mdqt_vel_dist(h, pv.data()));
        CKR(mdqt_populations_rows(h, pops_rows.data()));
        std::copy(pops_rows.begin(), pops_rows.end(), pops.begin() + (size_t)p.row0 * 3);
        CKR(mdqt_download_rows(h, NULL, V.data(), NULL, NULL, ld));
        bar.wait();
        if (g == 0) {
          OutputJob oj;
          oj.dir = dir; oj.counter = counter; oj.N = N; oj.d = d; oj.Epot0 = Epot0;
          oj.pvel = pv; oj.pops = pops; oj.vx.assign(V.begin(), V.begin() + N);
          push_output(writers, job, std::move(oj));
          counter++;
        }
        bar.wait();
        no++;
      }
      if (do_forces && !do_output && n == p.substeps_per_md) {
        int k = 1;
        for (;;) {
          int c0b = c0, tscb = tsc, o2, f2;
          double tb = t;
          int n2 = mdqt_schedule_next(&c0b, &tscb, &tb, p.substeps_per_md, o.sampleFreq, p.dtq, o.tmax, &o2, &f2);
          if (n2 != p.substeps_per_md || !f2 || o2) break;
          c0 = c0b; tsc = tscb; t = tb; k++;
        }
        CKR(mdqt_md_steps(h, k));
        nf += k; nsub += (long)k * n;
        continue;
      }
      // a partial MD step (around output()): forces over the own rows, n fused substeps, then the position all-gather
      if (do_forces) { CKR(mdqt_forces(h)); nf++; }
      CKR(mdqt_substeps(h, n));
      CKR(mdqt_comm_exchange_positions(h));
      nsub += n;
    }
    CKR(mdqt_download_rows(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
    bar.wait();
    if (g == 0) {
      c0_end = c0; t_end = t; nsub_total = nsub; nforce = nf; nout = no;
      wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    }
    mdqt_destroy(h);
#undef CKR
  };
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++) th.emplace_back(worker, g);
  for (auto& x : th) x.join();
  for (int g = 0; g < G; g++) if (rc[g]) return 1;
  for (auto& w : writers) w->finish();
  if (mdqt_io_write_conditions(dir, c0_end, N, counter, R.data(), V.data(), psi.data(), ld, vholder.data())) {
    fprintf(stderr, "mdqt_run: cannot write restart files into %s\n", dir);
    return 1;
  }
  if (!o.quiet)
    fprintf(stderr, "mdqt_run: job %u, N=%d row-decomposed over %d GPUs (%d rows each), t=%.6f, c0=%d: %ld substeps, %ld force calls, %ld outputs in %.3f s "
                    "(%.3e ion-steps/s, %.3e pair-interactions/s); files in %s\n", job, N, G, rows, t_end, c0_end, nsub_total, nforce, nout, wall,
            (double)N * nsub_total / wall, (double)N * N * nforce / wall, dir);
  return 0;
}

static void usage() {
  fprintf(stderr, "usage: mdqt_run [--program su|md|fz408l|mc408l|mc422l] <job> | --jobs a-b [--batch n] [--gpus n]\n"
                  "       [--Ge x] [--density x] [--sig0 x] [--Te x] [--fracOfSig x] [--N0 n] [--detuning x] [--detuningDP x] [--Om x]\n"
                  "       [--OmDP x] [--saveDirectory dir/] [--newRun 0|1] [--c0 n] [--tmax x] [--reNormalizewvFns 0|1] [--sampleFreq n]\n"
                  "       [--seed n] [--device n] [--writers n] [--fast-single] [--quiet]\n");
}

int mdqt_program_md(int argc, char** argv);      // mdqt_programs.cpp
int mdqt_program_fz408l(int argc, char** argv);
int mdqt_program_fz408q(int argc, char** argv);
int mdqt_program_fz422l(int argc, char** argv);
int mdqt_program_ts(int argc, char** argv);
int mdqt_program_mc408l(int argc, char** argv);
int mdqt_program_mc422l(int argc, char** argv);

int main(int argc, char** argv) {
  if (argc < 2) { usage(); return 2; }
  // mdqt_run --program md|fz408l|fz408q|fz422l|mc408l|mc422l|ts <job> [options]: the host loops of the reference's other programs (mdqt_programs.cpp)
  if (argc >= 3 && !strcmp(argv[1], "--program")) {
    std::vector<char*> av;
    av.push_back(argv[0]);
    for (int i = 3; i < argc; i++) av.push_back(argv[i]);
    if (!strcmp(argv[2], "md")) return mdqt_program_md((int)av.size(), av.data());
    if (!strcmp(argv[2], "fz408l")) return mdqt_program_fz408l((int)av.size(), av.data());
    if (!strcmp(argv[2], "fz408q")) return mdqt_program_fz408q((int)av.size(), av.data());
    if (!strcmp(argv[2], "fz422l")) return mdqt_program_fz422l((int)av.size(), av.data());
    if (!strcmp(argv[2], "ts")) return mdqt_program_ts((int)av.size(), av.data());
    if (!strcmp(argv[2], "mc408l")) return mdqt_program_mc408l((int)av.size(), av.data());
    if (!strcmp(argv[2], "mc422l")) return mdqt_program_mc422l((int)av.size(), av.data());
    if (strcmp(argv[2], "su")) { fprintf(stderr, "mdqt_run: unknown program %s (su, md, fz408l, fz408q, fz422l, mc408l, mc422l, ts)\n", argv[2]); return 2; }
    argc = (int)av.size();
    static std::vector<char*> keep;
    keep = av;
    argv = keep.data();
  }
  // defaults = the reference's globals (SU:56-78)
  std::map<std::string, std::string> opt = {
      {"Ge", "0.1"}, {"density", "2"}, {"sig0", "4.0"}, {"Te", "19.0"}, {"fracOfSig", "0"}, {"N0", "3500"}, {"detuning", "-1"},
      {"detuningDP", "1"}, {"Om", "1"}, {"OmDP", "1"}, {"saveDirectory", "dataLaserCool/"}, {"newRun", "1"}, {"c0", "0"},
      {"tmax", "30"}, {"reNormalizewvFns", "0"}, {"sampleFreq", "40"}, {"seed", ""}, {"device", "0"}, {"jobs", ""},
      {"batch", "64"}, {"gpus", "1"}, {"writers", "0"}};
  Options o;
  long job_a = -1, job_b = -1;
  int i = 1;
  if (argv[1][0] != '-') { job_a = job_b = (long)(unsigned)atof(argv[1]); i = 2; }  // SU:1145
  for (; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--quiet") { o.quiet = true; continue; }
    if (a == "--fast-single") { o.fast_single = 1; continue; }
    if (a.rfind("--", 0) != 0 || !opt.count(a.substr(2)) || i + 1 >= argc) { fprintf(stderr, "mdqt_run: bad option %s\n", a.c_str()); usage(); return 2; }
    opt[a.substr(2)] = argv[++i];
  }
  if (!opt["jobs"].empty()) {
    if (sscanf(opt["jobs"].c_str(), "%ld-%ld", &job_a, &job_b) != 2 || job_a < 0 || job_b < job_a) { fprintf(stderr, "mdqt_run: --jobs wants a-b\n"); return 2; }
  }
  if (job_a < 0) { usage(); return 2; }
  o.Ge = atof(opt["Ge"].c_str()); o.density = atof(opt["density"].c_str()); o.sig0 = atof(opt["sig0"].c_str());
  o.Te = atof(opt["Te"].c_str()); o.fracOfSig = atof(opt["fracOfSig"].c_str()); o.detuning = atof(opt["detuning"].c_str());
  o.detuningDP = atof(opt["detuningDP"].c_str()); o.Om = atof(opt["Om"].c_str()); o.OmDP = atof(opt["OmDP"].c_str());
  o.tmax = atof(opt["tmax"].c_str());
  o.N0 = atoi(opt["N0"].c_str()); o.newRun = atoi(opt["newRun"].c_str()); o.sampleFreq = atoi(opt["sampleFreq"].c_str());
  o.c0 = atoi(opt["c0"].c_str()); o.renorm = atoi(opt["reNormalizewvFns"].c_str()); o.device = atoi(opt["device"].c_str());
  o.saveDirectory = opt["saveDirectory"];
  const bool array = !opt["jobs"].empty();
  // seed of job j: the reference's srand48(time(NULL) + job) (SU:1219). With --seed S a single run uses S itself, an array
  // uses S + job -- so `mdqt_run j --seed S+j` reproduces job j of `mdqt_run --jobs a-b --seed S` bit for bit
  const bool have_seed = !opt["seed"].empty();
  o.seed = have_seed ? atol(opt["seed"].c_str()) : (long)(unsigned)time(NULL);
  o.seed_add_job = array || !have_seed;
  const int batch = std::max(1, atoi(opt["batch"].c_str()));
  int gpus = std::max(1, atoi(opt["gpus"].c_str()));
  const int ndev = mdqt_device_count();
  if (ndev <= 0) { fprintf(stderr, "mdqt_run: no CUDA device (the engine has no CPU fallback)\n"); return 1; }
  if (array && gpus > ndev) gpus = ndev;
  const long njobs = job_b - job_a + 1;
  int nwriters = atoi(opt["writers"].c_str());
  if (nwriters <= 0) nwriters = (int)std::min<long>(std::max(1u, std::thread::hardware_concurrency()), array ? std::min<long>(3 * njobs, 48) : 4);
  std::vector<std::unique_ptr<OutputWriter>> writers;
  for (int w = 0; w < nwriters; w++) writers.emplace_back(new OutputWriter());

  auto wall0 = std::chrono::steady_clock::now();
  std::vector<BatchResult> results;
  std::mutex rm;
  int failed = 0;
  if (!array && gpus > 1) {
    failed = run_rows(o, (unsigned)job_a, std::min(gpus, ndev), writers);
  } else if (!array) {
    BatchResult r = run_batch(o, (unsigned)job_a, 1, o.device, writers);
    failed = !r.ok;
  } else {
    // contiguous blocks of jobs per GPU, one host thread per GPU, `batch` jobs per handle
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++) {
      const long lo = job_a + njobs * g / gpus, hi = job_a + njobs * (g + 1) / gpus;  // [lo, hi)
      if (lo >= hi) continue;
      th.emplace_back([&, g, lo, hi] {
        for (long j = lo; j < hi; j += batch) {
          const int B = (int)std::min<long>(batch, hi - j);
          BatchResult r = run_batch(o, (unsigned)j, B, (o.device + g) % ndev, writers);
          std::lock_guard<std::mutex> lk(rm);
          results.push_back(r);
          if (!r.ok) failed++;
        }
      });
    }
    for (auto& t : th) t.join();
  }
  for (auto& w : writers) w->finish();  // every output file is on disk before the process reports success
  if (array && !o.quiet) {
    double ion_sub = 0;
    for (const BatchResult& r : results) ion_sub += r.ions * r.nsub;
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    fprintf(stderr, "mdqt_run: %ld jobs on %d GPU(s), batch %d, %d writer thread(s): %.3f s wall, %.3e ion-steps/s aggregate\n", njobs, gpus,
            batch, nwriters, wall, ion_sub / wall);
  }
  return failed ? 1 : 0;
}
