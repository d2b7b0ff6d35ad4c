// edmddna.cpp — config-driven driver: the reference's master_Code step loop
// (/root/reference/src/simulation.cpp:23-62, src/master.cpp:55-109,388-419) on the batched B200
// engine.  Reads the reference's unmodified input deck (config.txt, .prm, .crd) and writes its
// energies / trajectory / restart files in the same formats.
//
//   edmddna [--config ./config.txt] [--gpus N] [--gpu D] [--random 1|0] [--time T] [--nsteps N]
//
// --gpus N   run on N GPUs of this box (one process; replaces `aprun -n 91`, edmddna.pbs:20): tetrads and
//            the pair list are sharded over the GPUs, which exchange through peer memory (edmd_group_*).
// --gpu D    first device ordinal (default 0).
// --random 0 switches the Langevin random terms off (the reference has no such switch);
// --time T   pins the time(NULL) term of the random seed (edmd.cpp:179) for reproducible runs.
//
// Output is off the step loop's critical path: at an output point the merged state is unpacked on the
// device and drains into one of two pinned host buffers on a side stream (edmd_snapshot_begin) while the
// next block of steps is already running; a writer thread waits for the copy, rebuilds the per-base-pair
// arrays Master keeps (master.cpp:357-382) and writes the files (io.cpp:292-383).  The files are
// byte-identical to the blocking version.
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "edmd.hpp"
#include "io.hpp"

extern "C" {
#include "../../include/edmd_b200.h"
}

static void check(int rc, const char* what) {
    if (rc) { std::cerr << ">>> ERROR: " << what << ": " << edmd_last_error() << std::endl; std::exit(1); }
}

namespace {

struct OutputJob { int ticket, istep; bool info, crd; };

// Two pinned snapshot buffers + the thread that turns them into files, in submission order.
class Writer {
public:
    Writer(IO* io, edmd_group* grp, long long state_doubles, int n)
        : io_(io), grp_(grp), n_(n), total_(state_doubles), stop_(false) {
        for (int k = 0; k < 2; k++) {
            crd_[k] = (double*)edmd_alloc_pinned(sizeof(double) * (size_t)total_);
            vel_[k] = (double*)edmd_alloc_pinned(sizeof(double) * (size_t)total_);
            obs_[k] = (double*)edmd_alloc_pinned(sizeof(double) * 4 * (size_t)n_);
            if (!crd_[k] || !vel_[k] || !obs_[k]) check(-2, "edmd_alloc_pinned");
            busy_[k] = false;
        }
        flat_crd_.resize(3 * (size_t)io->crd.total_Atoms);
        flat_vel_.resize(3 * (size_t)io->crd.total_Atoms);
        thread_ = std::thread(&Writer::run, this);
    }
    ~Writer() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        thread_.join();
        for (int k = 0; k < 2; k++) { edmd_free_pinned(crd_[k]); edmd_free_pinned(vel_[k]); edmd_free_pinned(obs_[k]); }
    }
    // called from the step loop: start the device -> host copy of the current state and queue the file work
    void submit(int istep, bool info, bool crd) {
        const int k = next_;
        next_ ^= 1;
        {   // the buffer's previous contents must be on disk before the copy engine overwrites them
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return !busy_[k]; });
            busy_[k] = true;
        }
        check(edmd_group_snapshot_begin(grp_, k, crd_[k], vel_[k], info ? obs_[k] : nullptr), "edmd_group_snapshot_begin");
        { std::lock_guard<std::mutex> lk(mu_); jobs_.push_back(OutputJob{k, istep, info, crd}); }
        cv_.notify_all();
    }
    void drain() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return jobs_.empty() && !busy_[0] && !busy_[1]; });
    }

private:
    void run() {
        for (;;) {
            OutputJob job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                job = jobs_.front();
            }
            check(edmd_group_snapshot_wait(grp_, job.ticket), "edmd_group_snapshot_wait");
            const int k = job.ticket;
            // flat per-base-pair arrays as Master keeps them (master.cpp:357-382): after the merge all copies
            // of a base pair agree, so a plain scatter in tetrad order reproduces them
            size_t off = 0;
            for (int t = 0; t < n_; t++) {
                const int n3 = 3 * io_->tetrad[t].num_Atoms;
                std::memcpy(&flat_crd_[io_->displs[t]], crd_[k] + off, sizeof(double) * n3);
                std::memcpy(&flat_vel_[io_->displs[t]], vel_[k] + off, sizeof(double) * n3);
                off += n3;
            }
            if (job.info) {                                                    // Master::write_Info
                // master.cpp:393-401: sequential sums in tetrad order, mean temperature
                const double* o = obs_[k];
                double e[4] = {0.0, 0.0, 0.0, 0.0};
                for (int t = 0; t < n_; t++) {
                    e[0] += o[t]; e[1] += o[n_ + 2 * (size_t)t]; e[2] += o[n_ + 2 * (size_t)t + 1]; e[3] += o[3 * (size_t)n_ + t];
                }
                e[3] /= n_;
                io_->write_Energies(job.istep + io_->ntsync, e);
                io_->write_Trajectory(job.istep, io_->displs[io_->crd.num_BP - 3], flat_crd_.data());
            }
            if (job.crd) io_->update_Crd_File(flat_vel_.data(), flat_crd_.data());   // Master::write_Crds
            {
                std::lock_guard<std::mutex> lk(mu_);
                jobs_.pop_front();
                busy_[k] = false;
            }
            cv_.notify_all();
        }
    }

    IO* io_;
    edmd_group* grp_;
    int n_;
    long long total_;
    double *crd_[2], *vel_[2], *obs_[2];
    bool busy_[2];
    int next_ = 0;
    std::vector<double> flat_crd_, flat_vel_;
    std::deque<OutputJob> jobs_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_;
    std::thread thread_;
};

}  // namespace

int main(int argc, char** argv) {
    IO io;
    EDMD edmd;
    int gpu = 0, gpus = 1, random_mode = 1, nsteps_override = -1;
    long seed_time = (long)std::time(0);
    for (int i = 1; i < argc; i++) {
        const bool more = i + 1 < argc;
        if (!std::strcmp(argv[i], "--config") && more) io.config_File = argv[++i];
        else if (!std::strcmp(argv[i], "--gpu") && more) gpu = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--gpus") && more) gpus = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--random") && more) random_mode = std::atoi(argv[++i]) ? 1 : 0;
        else if (!std::strcmp(argv[i], "--time") && more) seed_time = std::atol(argv[++i]);
        else if (!std::strcmp(argv[i], "--nsteps") && more) nsteps_override = std::atoi(argv[++i]);
        else { std::cerr << "usage: edmddna [--config F] [--gpus N] [--gpu D] [--random 0|1] [--time T] [--nsteps N]\n"; return 2; }
    }
    if (gpus < 1) { std::cerr << "--gpus must be >= 1\n"; return 2; }
    const clock_t start = std::clock();

    // Master::initialise, master.cpp:55-109
    io.read_Cofig(&edmd);
    io.read_Prm();
    io.read_Crd();
    io.initialise_Tetrad_Crds();
    if (nsteps_override >= 0) io.nsteps = nsteps_override;
    io.ntwt -= io.ntwt % io.ntsync; if (io.ntwt == 0) io.ntwt = 1;   // master.cpp:65-66
    io.ntpr -= io.ntpr % io.ntsync; if (io.ntpr == 0) io.ntpr = 1;
    const int n = io.prm.num_Tetrads;

    std::cout << std::endl << "Initialising simulation..." << std::endl << std::endl;
    std::cout << "Simulation parameters:" << std::endl;
    std::cout << ">>> Total number of iterations  : " << io.nsteps << std::endl;
    std::cout << ">>> Frequency of synchronization: " << io.ntsync << std::endl;
    std::cout << ">>> Frequency of writing energy & temperature, trajectory: " << io.ntwt << std::endl;
    std::cout << ">>> Frequency of updating the coordinates file: " << io.ntpr << std::endl << std::endl;
    std::cout << "DNA Information:" << std::endl;
    std::cout << ">>> The number of DNA Base Pairs  : " << io.crd.num_BP << std::endl;
    std::cout << ">>> The number of DNA Tetrads     : " << n << std::endl;
    std::cout << ">>> Total number of atoms in DNA  : " << io.crd.total_Atoms << std::endl << std::endl;

    edmd_params p = {edmd.dt, edmd.gamma, edmd.tautp, edmd.temperature, edmd.scaled,
                     edmd.mole_Cutoff, edmd.atom_Cutoff, edmd.mole_Least};
    std::vector<int> devices(gpus);
    for (int r = 0; r < gpus; r++) devices[r] = gpu + r;
    edmd_group* grp = 0;
    check(edmd_group_create(&p, gpus, devices.data(), &grp), "edmd_group_create");
    check(edmd_group_upload_tetrads(grp, reinterpret_cast<edmd_tetrad*>(io.tetrad), n, io.crd.BP_Atoms), "edmd_group_upload_tetrads");
    check(edmd_group_set_rng(grp, seed_time, /*rank of the single worker*/ 1, 0), "edmd_group_set_rng");
    std::cout << ">>> GPUs: " << gpus << std::endl;

    {
        Writer writer(&io, grp, edmd_state_doubles(edmd_group_engine(grp, 0)), n);
        std::cout << "Simulation starting..." << std::endl;
        for (int istep = 0; istep < io.nsteps; istep += io.ntsync) {          // simulation.cpp:34-55
            std::cout << "istep: " << istep << std::endl;
            long long npairs = 0;
            check(edmd_group_build_pairlist(grp, &npairs), "edmd_group_build_pairlist");   // generate_Pair_Lists
            check(edmd_group_step(grp, io.ntsync, random_mode), "edmd_group_step");        // ntsync x (forces, velocity, coordinate)
            check(edmd_group_merge(grp), "edmd_group_merge");                              // merge_Vels_n_Crds
            const bool want_info = istep % io.ntwt == 0, want_crd = istep % io.ntpr == 0;
            if (want_info || want_crd) writer.submit(istep, want_info, want_crd);           // write_Info / write_Crds
        }
        check(edmd_group_sync(grp), "edmd_group_sync");
        writer.drain();
    }
    edmd_group_destroy(grp);
    std::cout << "Simulation ended.\nTime usage of simualtion: " << double(std::clock() - start) / CLOCKS_PER_SEC
              << std::endl << std::endl;
    return 0;
}
