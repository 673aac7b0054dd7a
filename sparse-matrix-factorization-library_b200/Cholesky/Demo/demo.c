/* Cholesky/Demo/demo.c -- runs the drop-in driver on the MatrixMarket files given on the command line
 * (reference: Cholesky/Demo/demo.c:4-27). */
#include <stdio.h>
#include "SparseFrame.h"

int main(int argc, char **argv)
{
    printf("SparseFrame Cholesky demo (B200 numeric engine)\n\n");
    printf("Size of Long = %zu, Float = %zu\n\n", sizeof(Long), sizeof(Float));
    return SparseFrame(argc, argv);
}
