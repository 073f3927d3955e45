/* TEST INFRASTRUCTURE: stand-in for <R_ext/Rdynload.h> */
#ifndef FAKE_RDYNLOAD_H
#define FAKE_RDYNLOAD_H
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct fake_dllinfo { const R_CallMethodDef* call_methods; int dynamic_symbols; } DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
#endif
