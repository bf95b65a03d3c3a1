/*
 * oracle/ref_access.c -- flat<->jagged copy helpers linked into oracle/_ref/liblpmc_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference keeps every array as row-by-row malloc'ed
 * T** / T*** (lpm_basic.c:380-445); Python (oracle/ref.py) uses these helpers to move
 * whole arrays in one call instead of one ctypes access per element.
 */
#include <stddef.h>
#include <string.h>

void lpmb_ref_gather_d2(double **src, int rows, int cols, double *dst)
{
    for (int i = 0; i < rows; i++)
        memcpy(dst + (size_t)i * cols, src[i], sizeof(double) * cols);
}
void lpmb_ref_scatter_d2(double **dst, int rows, int cols, const double *src)
{
    for (int i = 0; i < rows; i++)
        memcpy(dst[i], src + (size_t)i * cols, sizeof(double) * cols);
}
void lpmb_ref_gather_i2(int **src, int rows, int cols, int *dst)
{
    for (int i = 0; i < rows; i++)
        memcpy(dst + (size_t)i * cols, src[i], sizeof(int) * cols);
}
void lpmb_ref_scatter_i2(int **dst, int rows, int cols, const int *src)
{
    for (int i = 0; i < rows; i++)
        memcpy(dst[i], src + (size_t)i * cols, sizeof(int) * cols);
}
/* T*** [rows][cols][depth] <-> flat [rows][cols][depth] */
void lpmb_ref_gather_d3(double ***src, int rows, int cols, int depth, double *dst)
{
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            memcpy(dst + ((size_t)i * cols + j) * depth, src[i][j], sizeof(double) * depth);
}
void lpmb_ref_scatter_d3(double ***dst, int rows, int cols, int depth, const double *src)
{
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            memcpy(dst[i][j], src + ((size_t)i * cols + j) * depth, sizeof(double) * depth);
}
