// bch_coder.hpp -- host-side mirror of the reference's free-function toolbox
// (reference headers/bchCoder.h:10-48; definitions src/bchCoder.cpp).  Same names, argument
// meaning and ownership rules (functions that returned new[] memory still do), so code written
// against the reference header links against this one.  The GF(2)[x] helpers are set-up code
// (they run once per code); the per-frame Monte-Carlo work goes through the C ABI
// (include/pk_capi.h) to the sm_100a kernels -- see kaneko_processor.hpp and monte_carlo.hpp.
#pragma once
#include <cstdint>
#include <fstream>

// minimal polynomial of alpha^i, coefficients low -> high, *size = degree + 1 (bchCoder.cpp:25)
void findMinimalPolynomial(int i, int power, const unsigned long *fieldElements, int *size, unsigned char *res);
bool comparePoly(const unsigned char *poly1, int size1, const unsigned char *poly2, int size2);          // :92
// GF(2) product; the 4-argument form returns new[] memory owned by the caller (:104, :120)
unsigned char *multiplyPolynomials(const unsigned char *first, int size1, const unsigned char *second, int size2,
                                   int *sizeRes = nullptr);
void multiplyPolynomials(const unsigned char *first, int size1, const unsigned char *second, int size2,
                         unsigned char *res, int *sizeRes = nullptr);
// quotient (needRemainder = false) or remainder (true) as new[] memory (:134)
unsigned char *dividePolynomial(const unsigned char *first, int size1, const unsigned char *second, int size2,
                                int *size, bool needRemainder);
unsigned char *lcm(const unsigned char *first, int size1, const unsigned char *second, int size2, int *sizeRes);  // :217

// single-frame helpers of driver modes 4/5 (main.cpp:100-172); std::default_random_engine like the
// reference (:20-22).  The Monte-Carlo sweep does NOT use these: frames are drawn on the device.
unsigned char *generateRandomPoly(long k);                                                                // :228
void generateRandomPoly(unsigned char *res, long k);                                                      // :236
void addNoise(double standartDeviation, const unsigned char *codeword, double *wordWithNoise, unsigned long n);  // :243

void printVec(const unsigned char *poly, int size);                                                       // :261
void printVec(const unsigned long *poly, int size);
void printVec(const double *poly, int size);
void printVec(std::ofstream &out, const unsigned char *poly, int size);
void printVec(std::ofstream &out, const double *poly, int size);
void printMatrix(unsigned char **const matrix, int sizeI, int sizeJ = -1);
void printMatrix(std::ofstream &out, unsigned char **const matrix, int sizeI, int sizeJ = -1);
// n x n nested-BCH polarisation kernel (:317)
void makeMatrix(int power, const unsigned long *fieldElements, unsigned char **matrix);
