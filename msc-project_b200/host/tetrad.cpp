#include "tetrad.hpp"

// Sizes follow tetrad.cpp:23-36 of the reference; unlike the reference the arrays start zeroed.
void Tetrad::allocate_Tetrad_Arrays(void) {
    const int n3 = 3 * num_Atoms;
    avg = new double[n3]();
    masses = new double[n3]();
    abq = new double[n3]();
    eigenvalues = new double[num_Evecs > 0 ? num_Evecs : 1]();
    eigenvectors = Array::allocate_2D_Double_Array(num_Evecs, n3);
    velocities = new double[n3]();
    coordinates = new double[n3]();
    ED_Forces = new double[n3 + 1]();
    random_Terms = new double[n3]();
    NB_Forces = new double[n3 + 2]();
    temperature = 0.0;
}

void Tetrad::deallocate_Tetrad_Arrays(void) {
    delete[] avg; delete[] masses; delete[] abq; delete[] eigenvalues;
    Array::deallocate_2D_Double_Array(eigenvectors);
    delete[] velocities; delete[] coordinates;
    delete[] ED_Forces; delete[] random_Terms; delete[] NB_Forces;
    avg = masses = abq = eigenvalues = velocities = coordinates = 0;
    ED_Forces = random_Terms = NB_Forces = 0;
    eigenvectors = 0;
}
