// Test fixture: Bayesian logistic regression with a hierarchical scale, vectorised sampling statements.
data {
  int<lower=1> N;
  int<lower=1> K;
  array[N] int<lower=0, upper=1> y;
  array[N * K] real X;          // row-major design matrix
  real<lower=0, upper=1> phi;
}
parameters {
  real alpha;
  vector[K] w;
  real<lower=0> tau;
  real<lower=-1, upper=2> rho;
}
model {
  real eta;
  tau ~ exponential(1.5);
  alpha ~ student_t(4, 0, 2.5);
  w ~ normal(0, tau);
  rho ~ uniform(-1, 2);
  for (n in 1:N) {
    eta = alpha + rho;
    for (k in 1:K) eta += w[k] * X[(n - 1) * K + k];
    target += phi * bernoulli_logit_lpmf(y[n] | eta);
  }
}
