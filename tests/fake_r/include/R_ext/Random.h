/* TEST INFRASTRUCTURE: stand-in for <R_ext/Random.h> */
#ifndef FAKE_RANDOM_H
#define FAKE_RANDOM_H
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
#endif
