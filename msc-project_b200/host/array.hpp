// array.hpp — contiguous 2-D arrays with a row-pointer table.
// Same interface and memory layout as the reference's class Array
// (/root/reference/src/array.hpp:11-53, array.cpp:24-66): rows * cols values in ONE block,
// array[i] pointing at row i, so array[0] is a dense row-major matrix.  The engine relies on
// that layout for Tetrad::eigenvectors.
#ifndef EDMD_B200_HOST_ARRAY_HPP
#define EDMD_B200_HOST_ARRAY_HPP

class Array {
public:
    static double** allocate_2D_Double_Array(int rows, int cols);
    static void deallocate_2D_Double_Array(double** array);
    static int** allocate_2D_Int_Array(int rows, int cols);
    static void deallocate_2D_Int_Array(int** array);
};

#endif
