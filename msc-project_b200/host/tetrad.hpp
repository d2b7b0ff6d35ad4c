// tetrad.hpp — per-tetrad state, source- and layout-compatible with the reference's class Tetrad
// (/root/reference/src/tetrad.hpp:32-82).  The data members below are, in this order, exactly
// what `edmd_tetrad` (include/edmd_b200.h) describes, so a Tetrad* is an edmd_tetrad*.
#ifndef EDMD_B200_HOST_TETRAD_HPP
#define EDMD_B200_HOST_TETRAD_HPP

#include "array.hpp"

class Tetrad {
public:
    int num_Atoms;          // atoms in the tetrad (4 base pairs)
    int num_Evecs;          // PCA eigenvectors kept
    double* avg;            // [3N] average structure
    double* masses;         // [3N] mass of each atom, replicated for x, y, z
    double* abq;            // [3N] non-bonded parameters (A, B, q) per atom
    double* eigenvalues;    // [num_Evecs]
    double** eigenvectors;  // [num_Evecs][3N], one contiguous block
    double* ED_Forces;      // [3N+1], last entry = ED energy
    double* random_Terms;   // [3N]
    double* NB_Forces;      // [3N+2], last two = NB energy, electrostatic energy
    double temperature;
    double* velocities;     // [3N]
    double* coordinates;    // [3N]

    void allocate_Tetrad_Arrays(void);
    void deallocate_Tetrad_Arrays(void);
};

#endif
