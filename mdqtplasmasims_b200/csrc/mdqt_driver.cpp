// mdqt_driver.cpp -> mdqt_run: the time loop of the reference's main() (laserCoolingPlusExpansionMDQTSpeedUp.cpp:1139-1383)
// on top of the C ABI: same user inputs (by name, as options instead of edit-and-recompile), same directory tree,
// same output and restart files, hot path on the GPU.
//
//   mdqt_run <job> [--Ge 0.1] [--density 2] [--sig0 4] [--Te 19] [--fracOfSig 0] [--N0 3500] [--detuning -1]
//            [--detuningDP 1] [--Om 1] [--OmDP 1] [--saveDirectory dataLaserCool/] [--newRun 1] [--c0 0] [--tmax 30]
//            [--reNormalizewvFns 0] [--sampleFreq 40] [--seed <time+job>] [--device 0] [--quiet]
#include "../../include/mdqt.h"
#include "../../include/mdqt_io.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// output() formats ~9500 text lines per call (3 x 2001 velocity bins + one line per ion, SU:958-1024): about 3 ms of
// fprintf, as long as the 40 MD steps between two calls take on the GPU. A single writer thread does the formatting in
// call order while the time loop keeps the GPU busy; the files are byte-for-byte what a synchronous writer produces.
struct OutputJob {
  std::string dir;
  unsigned counter; int N;
  mdqt_diag d; double Epot0;
  std::vector<double> pvel, pops, vx;  // vx = V[0][0..N): the only velocity row write_populations prints
};
class OutputWriter {
 public:
  OutputWriter() : th_([this] { run(); }) {}
  ~OutputWriter() { finish(); }
  void push(OutputJob&& j) {
    std::unique_lock<std::mutex> lk(m_);
    cv_space_.wait(lk, [this] { return q_.size() < 4; });  // bounded: at most 4 outputs in flight
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
      mdqt_io_append_energies(j.dir.c_str(), j.d.t, j.d.ekin_x, j.d.ekin_y, j.d.ekin_z, j.d.epot, j.Epot0, j.d.vx_avg);
      mdqt_io_write_vel_dist(j.dir.c_str(), j.counter, j.pvel.data(), j.d.vx_avg);
      mdqt_io_write_populations(j.dir.c_str(), j.counter, j.N, j.vx.data(), j.pops.data());
    }
  }
  std::mutex m_;
  std::condition_variable cv_work_, cv_space_;
  std::deque<OutputJob> q_;
  bool done_ = false;
  std::thread th_;
};

static void die(const char* what) {
  fprintf(stderr, "mdqt_run: %s: %s\n", what, mdqt_last_error());
  exit(1);
}
#define CK(call) do { if ((call) != 0) die(#call); } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: mdqt_run <job> [--Ge x] [--density x] [--sig0 x] [--Te x] [--fracOfSig x] [--N0 n] [--detuning x]\n"
                    "       [--detuningDP x] [--Om x] [--OmDP x] [--saveDirectory dir/] [--newRun 0|1] [--c0 n] [--tmax x]\n"
                    "       [--reNormalizewvFns 0|1] [--sampleFreq n] [--seed n] [--device n] [--quiet]\n");
    return 2;
  }
  // defaults = the reference's globals (SU:56-78)
  std::map<std::string, std::string> opt = {
      {"Ge", "0.1"}, {"density", "2"}, {"sig0", "4.0"}, {"Te", "19.0"}, {"fracOfSig", "0"}, {"N0", "3500"}, {"detuning", "-1"},
      {"detuningDP", "1"}, {"Om", "1"}, {"OmDP", "1"}, {"saveDirectory", "dataLaserCool/"}, {"newRun", "1"}, {"c0", "0"},
      {"tmax", "30"}, {"reNormalizewvFns", "0"}, {"sampleFreq", "40"}, {"seed", ""}, {"device", "0"}};
  bool quiet = false;
  unsigned job = (unsigned)atof(argv[1]);  // SU:1145
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--quiet") { quiet = true; continue; }
    if (a.rfind("--", 0) != 0 || !opt.count(a.substr(2)) || i + 1 >= argc) { fprintf(stderr, "mdqt_run: bad option %s\n", a.c_str()); return 2; }
    opt[a.substr(2)] = argv[++i];
  }
  const double Ge = atof(opt["Ge"].c_str()), density = atof(opt["density"].c_str()), sig0 = atof(opt["sig0"].c_str());
  const double Te = atof(opt["Te"].c_str()), fracOfSig = atof(opt["fracOfSig"].c_str()), detuning = atof(opt["detuning"].c_str());
  const double detuningDP = atof(opt["detuningDP"].c_str()), Om = atof(opt["Om"].c_str()), OmDP = atof(opt["OmDP"].c_str());
  const double tmax = atof(opt["tmax"].c_str());
  const int N0 = atoi(opt["N0"].c_str()), newRun = atoi(opt["newRun"].c_str()), sampleFreq = atoi(opt["sampleFreq"].c_str());
  int c0 = atoi(opt["c0"].c_str());
  const long seed = opt["seed"].empty() ? (long)((unsigned)time(NULL) + job) : atol(opt["seed"].c_str());  // SU:1219
  const int ld = N0 + 1000;  // SU:126

  char dir[1024];
  if (mdqt_io_dirname(dir, sizeof(dir), opt["saveDirectory"].c_str(), Ge, density, sig0, Te, fracOfSig, detuning, detuningDP, Om,
                      OmDP, N0, job, 1)) { fprintf(stderr, "mdqt_run: directory name too long\n"); return 1; }

  std::vector<double> R(3 * (size_t)ld), V(3 * (size_t)ld), psi((size_t)ld * 24), tPart(ld, 0.0), vholder((size_t)3 * MDQT_NUM_VINTERVALS * ld, 0.0);
  std::vector<double> pvel(3 * 2001), pops((size_t)ld * 3);
  unsigned counter = 0;
  double t = 0.0, L = 0, lDeb = 0;
  int N;
  if (newRun == 1) {
    N = mdqt_io_init_su(seed, N0, Ge, ld, R.data(), V.data(), psi.data(), tPart.data(), &L, &lDeb);  // init(), SU:289-348
    if (N < 0) { fprintf(stderr, "mdqt_run: more than N0+1000 ions drawn\n"); return 1; }
    printf("%i\n", N);  // SU:338
    c0 = -1;            // SU:347
  } else {
    N = mdqt_io_read_conditions(dir, c0, ld, R.data(), V.data(), psi.data(), &counter, &t, vholder.data());  // SU:785-916
    if (N < 0) { fprintf(stderr, "mdqt_run: cannot read restart files for c0=%d in %s (%d)\n", c0, dir, N); return 1; }
  }

  mdqt_params p;
  CK(mdqt_params_su(&p, Ge, density, sig0, Te, fracOfSig, detuning, detuningDP, Om, OmDP, N0, N));
  p.traj0 = (int)job; p.seed = (uint64_t)seed; p.device = atoi(opt["device"].c_str());
  p.renormalize = atoi(opt["reNormalizewvFns"].c_str());
  mdqt_handle* h = NULL;
  CK(mdqt_create(&p, &h));
  CK(mdqt_upload_state(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
  // RNG substep counter: continue the stream where a previous run of this job stopped
  uint64_t sub0 = newRun == 1 ? 0 : (uint64_t)llround(t / p.dtq);
  CK(mdqt_set_time(h, t, sub0));

  double Epot0 = 0.0;
  CK(mdqt_epot(h, &Epot0));  // Epotential(); Epot0 = Epot (SU:345-346). On resume the reference leaves Epot0 = 0 (Q8):
  if (newRun != 1) Epot0 = 0.0;

  OutputWriter writer;
  int tsc = p.substeps_per_md;  // timeStepCounter (SU:1235)
  auto wall0 = std::chrono::steady_clock::now();
  long nsub_total = 0, nforce = 0, nout = 0;
  for (;;) {
    int do_output, do_forces;
    int n = mdqt_schedule_next(&c0, &tsc, &t, p.substeps_per_md, sampleFreq, p.dtq, tmax, &do_output, &do_forces);
    if (n == 0) break;
    if (do_output) {  // output(), SU:917-1032
      mdqt_diag d;
      CK(mdqt_diagnostics(h, &d));
      CK(mdqt_vel_dist(h, pvel.data()));
      CK(mdqt_populations(h, pops.data()));
      CK(mdqt_download_state(h, NULL, V.data(), NULL, NULL, ld));
      OutputJob job;
      job.dir = dir; job.counter = counter; job.N = N; job.d = d; job.Epot0 = Epot0;
      job.pvel = pvel; job.pops.assign(pops.begin(), pops.begin() + (size_t)N * 3); job.vx.assign(V.begin(), V.begin() + N);
      writer.push(std::move(job));
      counter++;
      nout++;
    }
    if (do_forces && !do_output && n == p.substeps_per_md) {
      // whole MD steps with no output() in between: hand the run of them to mdqt_md_steps (one replayed CUDA graph)
      int k = 1;
      for (;;) {
        int c0b = c0, tscb = tsc, o2, f2;
        double tb = t;
        int n2 = mdqt_schedule_next(&c0b, &tscb, &tb, p.substeps_per_md, sampleFreq, p.dtq, tmax, &o2, &f2);
        if (n2 != p.substeps_per_md || !f2 || o2) break;
        c0 = c0b; tsc = tscb; t = tb; k++;
      }
      CK(mdqt_md_steps(h, k));
      nforce += k; nsub_total += (long)k * n;
      continue;
    }
    if (do_forces) { CK(mdqt_forces(h)); nforce++; }
    CK(mdqt_substeps(h, n));
    nsub_total += n;
  }
  writer.finish();  // every output file is on disk before the restart files are written
  CK(mdqt_download_state(h, R.data(), V.data(), psi.data(), tPart.data(), ld));
  double t_dev; uint64_t s_dev;
  CK(mdqt_get_time(h, &t_dev, &s_dev));
  if (t_dev != t) fprintf(stderr, "mdqt_run: warning: host/device clocks differ (%.17g vs %.17g)\n", t, t_dev);
  if (mdqt_io_write_conditions(dir, c0, N, counter, R.data(), V.data(), psi.data(), ld, vholder.data())) {  // SU:1381
    fprintf(stderr, "mdqt_run: cannot write restart files into %s\n", dir);
    return 1;
  }
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
  if (!quiet)
    fprintf(stderr, "mdqt_run: job %u, N=%d, t=%.6f, c0=%d: %ld substeps, %ld force calls, %ld outputs in %.3f s (%.3e ion-steps/s); files in %s\n",
            job, N, t, c0, nsub_total, nforce, nout, wall, (double)N * nsub_total / wall, dir);
  mdqt_destroy(h);
  return 0;
}
