# .Call stub of the sampler: the reference generates this file with Rcpp::compileAttributes() (R/RcppExports.R:4-6 there);
# the replacement DLL exports the same registered routine, so R/gpirtMCMC.R of the reference runs unchanged.
.gpirtMCMC <- function(y, theta, sample_iterations, burn_iterations, beta_prior_means, beta_prior_sds, beta_step_sizes) {
    .Call(`_gpirt_gpirtMCMC`, y, theta, sample_iterations, burn_iterations, beta_prior_means, beta_prior_sds, beta_step_sizes)
}
