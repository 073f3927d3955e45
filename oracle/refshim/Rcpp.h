/* TEST INFRASTRUCTURE — see RcppArmadillo.h in this directory. */
#include "RcppArmadillo.h"
