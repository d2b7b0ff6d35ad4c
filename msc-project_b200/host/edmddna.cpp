// edmddna.cpp — config-driven driver: the reference's master_Code step loop
// (/root/reference/src/simulation.cpp:23-62, src/master.cpp:55-109,388-419) on the batched B200
// engine.  Reads the reference's unmodified input deck (config.txt, .prm, .crd) and writes its
// energies / trajectory / restart files in the same formats.
//
//   edmddna [--config ./config.txt] [--gpu 0] [--random 1|0] [--time T] [--nsteps N]
//
// --random 0 switches the Langevin random terms off (the reference has no such switch);
// --time T pins the time(NULL) term of the random seed (edmd.cpp:179) for reproducible runs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <vector>

#include "edmd.hpp"
#include "io.hpp"

extern "C" {
#include "../../include/edmd_b200.h"
}

static void check(int rc, const char* what) {
    if (rc) { std::cerr << ">>> ERROR: " << what << ": " << edmd_last_error() << std::endl; std::exit(1); }
}

int main(int argc, char** argv) {
    IO io;
    EDMD edmd;
    int gpu = 0, random_mode = 1, nsteps_override = -1;
    long seed_time = (long)std::time(0);
    for (int i = 1; i < argc; i++) {
        const bool more = i + 1 < argc;
        if (!std::strcmp(argv[i], "--config") && more) io.config_File = argv[++i];
        else if (!std::strcmp(argv[i], "--gpu") && more) gpu = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--random") && more) random_mode = std::atoi(argv[++i]) ? 1 : 0;
        else if (!std::strcmp(argv[i], "--time") && more) seed_time = std::atol(argv[++i]);
        else if (!std::strcmp(argv[i], "--nsteps") && more) nsteps_override = std::atoi(argv[++i]);
        else { std::cerr << "usage: edmddna [--config F] [--gpu D] [--random 0|1] [--time T] [--nsteps N]\n"; return 2; }
    }
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
    edmd_engine* eng = 0;
    check(edmd_create(&p, gpu, &eng), "edmd_create");
    check(edmd_upload_tetrads(eng, reinterpret_cast<edmd_tetrad*>(io.tetrad), n, io.crd.BP_Atoms), "edmd_upload_tetrads");
    check(edmd_set_rng(eng, seed_time, /*rank of the single worker*/ 1, 0), "edmd_set_rng");

    std::vector<double> coordinates(3 * (size_t)io.crd.total_Atoms), velocities(3 * (size_t)io.crd.total_Atoms);
    std::cout << "Simulation starting..." << std::endl;

    for (int istep = 0; istep < io.nsteps; istep += io.ntsync) {          // simulation.cpp:34-55
        std::cout << "istep: " << istep << std::endl;
        long long npairs = 0;
        check(edmd_build_pairlist(eng, &npairs), "edmd_build_pairlist");   // generate_Pair_Lists
        check(edmd_step(eng, io.ntsync, random_mode), "edmd_step");        // ntsync x (forces, velocity, coordinate)
        check(edmd_merge(eng), "edmd_merge");                              // merge_Vels_n_Crds
        const bool want_info = istep % io.ntwt == 0, want_crd = istep % io.ntpr == 0;
        if (want_info || want_crd) {
            check(edmd_get_state(eng, reinterpret_cast<edmd_tetrad*>(io.tetrad), n), "edmd_get_state");
            // flat per-base-pair arrays as Master keeps them (master.cpp:357-382): after the merge
            // all copies of a base pair agree, so a plain scatter reproduces them
            for (int t = 0; t < n; t++)
                for (int k = 0; k < 3 * io.tetrad[t].num_Atoms; k++) {
                    coordinates[io.displs[t] + k] = io.tetrad[t].coordinates[k];
                    velocities[io.displs[t] + k] = io.tetrad[t].velocities[k];
                }
        }
        if (want_info) {                                                   // Master::write_Info
            double e[4];
            check(edmd_get_energies(eng, e), "edmd_get_energies");
            io.write_Energies(istep + io.ntsync, e);
            io.write_Trajectory(istep, io.displs[io.crd.num_BP - 3], coordinates.data());
        }
        if (want_crd) io.update_Crd_File(velocities.data(), coordinates.data());   // Master::write_Crds
    }
    check(edmd_sync(eng), "edmd_sync");
    edmd_destroy(eng);
    std::cout << "Simulation ended.\nTime usage of simualtion: " << double(std::clock() - start) / CLOCKS_PER_SEC
              << std::endl << std::endl;
    return 0;
}
