/* TEST INFRASTRUCTURE: stand-in for <R_ext/Utils.h> */
#ifndef FAKE_UTILS_H
#define FAKE_UTILS_H
void R_CheckUserInterrupt(void);
int R_ToplevelExec(void (*fun)(void*), void* data);
#endif
