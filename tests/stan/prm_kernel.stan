// Test fixture (written for this repo): the density of the built-in "PRMwCD" device model -- Poisson regression on
// precomputed Gaussian-kernel features with an exponential-power prior of scale Gamma on the coefficients -- phrased with
// a log-rate accumulator and poisson_log.  Reads the same data file as the built-in model.
data {
  int<lower=1> N;
  int<lower=1> M;
  real<lower=0> q;
  int<lower=1> Clength;
  array[N] int<lower=0> y;
  array[N * Clength] real Xkernel;    // row-major N x Clength
  real<lower=0, upper=1> phi;
}
parameters {
  array[M] real Beta;
  real<lower=0> Gamma;
}
model {
  target += inv_gamma_lpdf(Gamma | 2, 1.3);
  for (i in 1:N) {
    real eta = Beta[1];
    for (j in 1:Clength)
      eta += Beta[j + 1] * Xkernel[(i - 1) * Clength + j];
    target += phi * poisson_log_lpmf(y[i] | eta);
  }
  for (i in 2:M)
    target += -log(Gamma) - pow(fabs(Beta[i] / Gamma), q);
}
