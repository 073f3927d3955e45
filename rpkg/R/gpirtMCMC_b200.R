#' gpirtMCMC() with the draw-storage options of the B200 back end
#'
#' Same arguments and defaults as \code{gpirtMCMC()} of the reference package, plus
#' \describe{
#'   \item{thin}{keep the draws of every \code{thin}-th sampling iteration only; \code{theta}, \code{beta} and \code{f}
#'     then hold \code{1 + sample_iterations \%/\% thin} slots (slot 1 = initial values). IRFs still average over all
#'     sampling iterations.}
#'   \item{store_f}{\code{FALSE}: do not return the \code{n x m x slots} array of f draws (\code{f} is \code{NULL}).}
#'   \item{f_summary}{\code{TRUE}: also return \code{f_mean} and \code{f_sd}, the posterior mean and standard deviation of
#'     f over all sampling iterations, accumulated on the GPU.}
#' }
#' @export
gpirtMCMC_b200 <- function(data, sample_iterations, burn_iterations,
                           vote_codes = list(yea = 1:3, nay = 4:6, missing = c(0, 7:9, NA)),
                           beta_prior_means = matrix(0, nrow = 2, ncol = ncol(data)),
                           beta_prior_sds = matrix(3, nrow = 2, ncol = ncol(data)),
                           beta_proposal_sds = matrix(0.1, nrow = 2, ncol = ncol(data)),
                           theta_init = NULL, thin = 1L, store_f = TRUE, f_summary = FALSE) {
    data <- as.response_matrix(data, vote_codes)
    if (is.null(theta_init)) theta_init <- rnorm(nrow(data))
    .Call(`_gpirt_gpirtMCMC_b200`, data, theta_init, sample_iterations, burn_iterations, beta_prior_means,
          beta_prior_sds, beta_proposal_sds, as.integer(thin), as.integer(store_f), as.integer(f_summary))
}
