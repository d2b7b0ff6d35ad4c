#include "array.hpp"

namespace {
template <class T>
T** make_rows(int rows, int cols) {
    if (rows < 1) rows = 1;
    T** table = new T*[rows];
    T* block = new T[(long)rows * (cols > 0 ? cols : 1)];
    for (int r = 0; r < rows; r++) table[r] = block + (long)r * cols;
    return table;
}
template <class T>
void drop_rows(T** table) {
    if (!table) return;
    delete[] table[0];
    delete[] table;
}
}  // namespace

double** Array::allocate_2D_Double_Array(int rows, int cols) { return make_rows<double>(rows, cols); }
void Array::deallocate_2D_Double_Array(double** array) { drop_rows(array); }
int** Array::allocate_2D_Int_Array(int rows, int cols) { return make_rows<int>(rows, cols); }
void Array::deallocate_2D_Int_Array(int** array) { drop_rows(array); }
