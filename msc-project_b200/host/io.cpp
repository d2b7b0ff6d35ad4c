// io.cpp — see io.hpp.  Formats restated from the reference (file:line cited per function).
#include "io.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace {
[[noreturn]] void fatal(const char* msg) {
    std::cout << ">>> ERROR: " << msg << std::endl;   // same wording/exit code as io.cpp:107-108
    std::exit(1);
}
// third whitespace-separated token of a `key = value` line (the key text itself is ignored,
// io.cpp:78-100)
std::string value_of(const std::string& line) {
    std::istringstream in(line);
    std::string key, eq, val;
    in >> key >> eq >> val;
    return val;
}
template <class T>
void parse(const std::string& text, T& out) {
    std::istringstream in(text);
    in >> out;
}
void put_fixed(std::ostream& out, double v) {   // io.cpp:330,364: fixed, width 10, 4 decimals, one blank
    out << std::fixed << std::setw(10) << std::setprecision(4) << v << " ";
}
}  // namespace

IO::IO(void) {   // defaults of io.cpp:23-39
    crd.num_BP = 0; crd.BP_Atoms = 0; crd.total_Atoms = 0; crd.BP_Vels = 0; crd.BP_Crds = 0;
    prm.num_Tetrads = 0;
    tetrad = 0; displs = 0;
    irest = 0; nsteps = 100000; ntsync = 100; ntwt = 100; ntpr = 1000;
    prm_File = "./test/GC90c12.prm";
    crd_File = "./test/GC90_6c.crd";
    energy_File = "./data/energies.eng";
    trj_File = "./data/trajectory.trj";
    new_Crd_File = "./data/crd.crd";
    config_File = "./config.txt";
}

IO::~IO(void) {
    delete[] displs;
    delete[] crd.BP_Atoms; delete[] crd.BP_Crds; delete[] crd.BP_Vels;
    if (tetrad) {
        for (int i = 0; i < prm.num_Tetrads; i++) tetrad[i].deallocate_Tetrad_Arrays();
        delete[] tetrad;
    }
}

// io.cpp:60-116: 17 lines in fixed order; time-like values are converted to internal units
void IO::read_Cofig(EDMD* edmd) {
    std::ifstream fin(config_File.c_str());
    if (!fin.is_open()) fatal("Can not open the Config file!");
    std::string line;
    for (int k = 1; k <= 17 && std::getline(fin, line); k++) {
        if (line.size() > 99) line.resize(99);   // the reference reads into a 100-char buffer
        const std::string v = value_of(line);
        switch (k) {
            case 1: parse(v, irest); break;
            case 2: parse(v, nsteps); break;
            case 3: parse(v, ntsync); break;
            case 4: parse(v, ntwt); break;
            case 5: parse(v, ntpr); break;
            case 6: parse(v, edmd->dt); edmd->dt *= edmd->constants.timefac; break;
            case 7: parse(v, edmd->gamma); edmd->gamma /= edmd->constants.timefac; break;
            case 8: parse(v, edmd->tautp); edmd->tautp *= edmd->constants.timefac; break;
            case 9: parse(v, edmd->temperature); edmd->scaled = edmd->constants.Boltzmann * edmd->temperature; break;
            case 10: parse(v, edmd->mole_Cutoff); break;
            case 11: parse(v, edmd->atom_Cutoff); break;
            case 12: parse(v, edmd->mole_Least); break;
            case 13: prm_File = v; break;
            case 14: crd_File = v; break;
            case 15: energy_File = v; break;
            case 16: trj_File = v; break;
            case 17: new_Crd_File = v; break;
        }
    }
    // io.cpp:112-114: outputs of an earlier run are removed.  Deliberate deviation: the reference
    // also removes new_Crd_File when irest != 0 and then fails to open it in read_Crd
    // (io.cpp:193,228) — its restart path cannot work as shipped — so a restart keeps that file.
    std::remove(energy_File.c_str());
    std::remove(trj_File.c_str());
    if (irest == 0) std::remove(new_Crd_File.c_str());
}

// io.cpp:120-181: count; per tetrad: atoms evecs | 3N avg | N masses | 3N (A,B,q) | per evec: value + 3N
void IO::read_Prm(void) {
    std::ifstream fin(prm_File.c_str());
    if (!fin.is_open()) fatal("Can not open the prm file!");
    fin >> prm.num_Tetrads;
    tetrad = new Tetrad[prm.num_Tetrads];
    for (int t = 0; t < prm.num_Tetrads; t++) {
        Tetrad& T = tetrad[t];
        fin >> T.num_Atoms >> T.num_Evecs;
        T.allocate_Tetrad_Arrays();
        const int n3 = 3 * T.num_Atoms;
        for (int k = 0; k < n3; k++) fin >> T.avg[k];
        for (int a = 0; a < T.num_Atoms; a++) {      // one mass per atom, spread over x, y, z
            double m; fin >> m;
            T.masses[3 * a] = T.masses[3 * a + 1] = T.masses[3 * a + 2] = m;
        }
        for (int k = 0; k < n3; k++) fin >> T.abq[k];
        for (int j = 0; j < T.num_Evecs; j++) {
            fin >> T.eigenvalues[j];
            for (int k = 0; k < n3; k++) fin >> T.eigenvectors[j][k];
        }
    }
}

// io.cpp:185-232: num_BP | atoms per bp | per bp: 3n coordinates (+ 3n velocities when restarting)
void IO::read_Crd(void) {
    std::ifstream fin((irest == 0 ? crd_File : new_Crd_File).c_str());
    if (!fin.is_open()) fatal("Can not open the crd file!");
    fin >> crd.num_BP;
    crd.BP_Atoms = new int[crd.num_BP];
    crd.total_Atoms = 0;
    for (int b = 0; b < crd.num_BP; b++) { fin >> crd.BP_Atoms[b]; crd.total_Atoms += crd.BP_Atoms[b]; }
    crd.BP_Crds = new double[3 * crd.total_Atoms]();
    crd.BP_Vels = new double[3 * crd.total_Atoms]();
    generate_Displacements();
    for (int b = 0; b < crd.num_BP; b++) {
        for (int k = displs[b]; k < displs[b + 1]; k++) fin >> crd.BP_Crds[k];
        if (irest != 0)
            for (int k = displs[b]; k < displs[b + 1]; k++) fin >> crd.BP_Vels[k];
    }
}

// io.cpp:236-247
void IO::generate_Displacements(void) {
    delete[] displs;
    displs = new int[crd.num_BP + 1];
    displs[0] = 0;
    for (int b = 0; b < crd.num_BP; b++) displs[b + 1] = displs[b] + 3 * crd.BP_Atoms[b];
}

// io.cpp:251-288: tetrad t = base pairs t .. t+3
void IO::initialise_Tetrad_Crds(void) {
    for (int t = 0; t < prm.num_Tetrads; t++) {
        Tetrad& T = tetrad[t];
        const int atoms = crd.BP_Atoms[t] + crd.BP_Atoms[t + 1] + crd.BP_Atoms[t + 2] + crd.BP_Atoms[t + 3];
        if (atoms != T.num_Atoms) {
            std::cout << ">>> ERROR: The number of atoms in crd file is wrong." << std::endl;
            std::abort();   // the reference calls MPI_Abort here (io.cpp:269)
        }
        const int first = displs[t];
        for (int k = 0; k < 3 * T.num_Atoms; k++) {
            T.velocities[k] = irest == 0 ? 0.0 : crd.BP_Vels[first + k];
            T.coordinates[k] = crd.BP_Crds[first + k];
            T.ED_Forces[k] = 0.0; T.random_Terms[k] = 0.0; T.NB_Forces[k] = 0.0;
        }
    }
}

// io.cpp:292-314: one appended line, default float format with 8 significant digits
void IO::write_Energies(int istep, double energies[]) {
    std::ofstream fout(energy_File.c_str(), std::ios_base::app);
    if (!fout.is_open()) fatal("Can not open the energy file!");
    fout << "Iteration, ED Energy, NB_Energy, ELE_Energy, Temperature: " << istep << ", ";
    fout << std::setprecision(8) << energies[0] << ", ";
    fout << std::setprecision(8) << energies[1] << ", ";
    fout << std::setprecision(8) << energies[2] << ", ";
    fout << std::setprecision(8) << energies[3] << std::endl;
}

// io.cpp:318-341: appended frame of `index` values, 10 per line; atom count before the first frame
void IO::write_Trajectory(int istep, int index, double* coordinates) {
    std::ofstream fout(trj_File.c_str(), std::ios_base::app);
    if (!fout.is_open()) fatal("Can not open the trajectory file!");
    if (istep == 0) fout << crd.total_Atoms << std::endl;
    for (int k = 0; k < index; k++) {
        put_fixed(fout, coordinates[k]);
        if ((k + 1) % 10 == 0) fout << std::endl;
    }
}

// io.cpp:345-383: restart file = .crd layout with velocities after each base pair's coordinates;
// line breaks follow the GLOBAL value index
void IO::update_Crd_File(double* velocities, double* coordinates) {
    std::ofstream fout(new_Crd_File.c_str(), std::ios_base::out);
    if (!fout.is_open()) fatal("Can not open the new crd file!");
    fout << crd.num_BP << std::endl;
    for (int b = 0; b < crd.num_BP; b++) fout << crd.BP_Atoms[b] << std::endl;
    for (int b = 0; b < crd.num_BP; b++) {
        for (int k = displs[b]; k < displs[b + 1]; k++) {
            put_fixed(fout, coordinates[k]);
            if ((k + 1) % 10 == 0) fout << std::endl;
        }
        fout << std::endl;
        for (int k = displs[b]; k < displs[b + 1]; k++) {
            put_fixed(fout, velocities[k]);
            if ((k + 1) % 10 == 0) fout << std::endl;
        }
        fout << std::endl;
    }
}
