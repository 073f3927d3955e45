/* TEST INFRASTRUCTURE: stand-in for <R.h> */
#ifndef FAKE_R_H
#define FAKE_R_H
#ifdef __cplusplus
extern "C" {
#endif
void Rprintf(const char*, ...);
typedef enum { FALSE = 0, TRUE } Rboolean;   /* R_ext/Boolean.h */
#ifdef __cplusplus
}
#endif
#endif
