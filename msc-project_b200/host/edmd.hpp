// edmd.hpp — the reference's class EDMD surface (/root/reference/src/edmd.hpp:54-170) on top of
// the B200 engine.  Same public fields, same five methods on Tetrad*, same units (internal units,
// i.e. after the conversion of io.cpp:85-91).  Each method is a batch-of-one GPU call through the
// C ABI (include/edmd_b200.h) — meant for call-by-call drop-in use and per-function parity tests;
// production runs drive the batched engine (see edmddna.cpp).  There is no CPU implementation
// behind this class: without a CUDA device every method aborts like the reference's fatal paths.
#ifndef EDMD_B200_HOST_EDMD_HPP
#define EDMD_B200_HOST_EDMD_HPP

#include "tetrad.hpp"

struct edmd_engine;

typedef struct _Constants {
    double Boltzmann;   // Boltzmann constant in internal units
    double timefac;     // time unit conversion factor
} Constants;

class EDMD {
public:
    Constants constants;
    double dt;            // timestep (internal units)
    double gamma;         // friction coefficient
    double tautp;         // Berendsen coupling time
    double temperature;   // K
    double scaled;        // kT in internal units
    double mole_Cutoff;   // tetrad-level cutoff (centroid distance)
    double atom_Cutoff;   // atom-level cutoff
    double mole_Least;    // minimum chain separation of interacting tetrads

    EDMD(void);
    ~EDMD(void);
    void initialise(double _dt, double _gamma, double _tautp, double _temperature, double _scaled,
                    double _mole_Cutoff, double _atom_Cutoff, double _mole_Least);
    void calculate_ED_Forces(Tetrad* tetrad);
    void calculate_Random_Terms(Tetrad* tetrad, int rank);
    void calculate_NB_Forces(Tetrad* t1, Tetrad* t2);
    void update_Velocities(Tetrad* tetrad);
    void update_Coordinates(Tetrad* tetrad);

    // GPU ordinal used by this object's engines (default 0); set before the first method call.
    int device;
    // When >= 0 the Langevin seed uses this instead of time(NULL) (reproducible runs / tests).
    long pinned_time;
    // calculate_Random_Terms calls made so far in this process (the reference's function-static
    // RNG_Seed - 13579, edmd.cpp:173); public so that a run can be replayed.
    static long random_calls;

private:
    edmd_engine* one;     // engine holding a single tetrad
    edmd_engine* two;     // engine holding a tetrad pair
    void sync_params(edmd_engine* e);
    edmd_engine* engine_for(Tetrad* t);
    edmd_engine* engine_for(Tetrad* t1, Tetrad* t2);
    EDMD(const EDMD&);
    EDMD& operator=(const EDMD&);
};

#endif
