// io.hpp — input decks and output files of the EDMD code, byte-compatible with the reference's
// class IO (/root/reference/src/io.hpp:36-204, io.cpp:60-383): config.txt (17 positional
// `key = value` lines), the tetrad parameter file (.prm), the coordinate/restart file (.crd) and
// the energies / trajectory / restart writers.  Host-side only; nothing here touches the GPU.
#ifndef EDMD_B200_HOST_IO_HPP
#define EDMD_B200_HOST_IO_HPP

#include <string>

#include "edmd.hpp"
#include "tetrad.hpp"

typedef struct _Crd {
    int num_BP;         // base pairs in the file (ring size + 3 repeated)
    int* BP_Atoms;      // atoms per base pair
    int total_Atoms;
    double* BP_Vels;    // [3*total_Atoms]
    double* BP_Crds;    // [3*total_Atoms]
} Crd;

typedef struct _Prm {
    int num_Tetrads;
} Prm;

class IO {
public:
    Crd crd;
    Prm prm;
    Tetrad* tetrad;     // num_Tetrads entries, arrays owned by this object
    int* displs;        // [num_BP+1] offset (in doubles) of each base pair in the flat arrays
    int irest;          // 0: start from crd_File, else restart from new_Crd_File (with velocities)
    int nsteps, ntsync, ntwt, ntpr;
    std::string prm_File, crd_File, energy_File, trj_File, new_Crd_File;
    // Not in the reference: where the configuration is read from (reference: always ./config.txt)
    std::string config_File;

    IO(void);
    ~IO(void);
    void read_Cofig(EDMD* edmd);
    void read_Prm(void);
    void read_Crd(void);
    void generate_Displacements(void);
    void initialise_Tetrad_Crds(void);
    void write_Energies(int istep, double energies[]);
    void write_Trajectory(int istep, int index, double* coordinates);
    void update_Crd_File(double* velocities, double* coordinates);
};

#endif
