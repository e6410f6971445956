/* Test fixture: linear regression the way the Stan manual writes it -- matrix * vector, transformed data,
   transformed parameters, a vector-valued declaration with an initialiser, generated quantities (ignored). */
data {
  int<lower=1> N;
  int<lower=1> K;
  matrix[N, K] X;
  vector[N] y;
  real<lower=0, upper=1> phi;
}
transformed data {
  int H = K - 1;                         // folded at generation time
  real prior_scale = 2.5 * 2;
  vector[N] w = rep_vector(0.25, N) + 0.75;
}
parameters {
  real alpha;
  vector[K] beta;
  real<lower=0> sigma;
}
transformed parameters {
  vector[N] mu = alpha + X * beta;
}
model {
  beta ~ std_normal();
  alpha ~ normal(0, prior_scale);
  sigma ~ exponential(1);
  for (h in 1:H)
    target += -0.5 * square(beta[h + 1] - beta[h]);      // random-walk smoothing of neighbouring coefficients
  target += phi * normal_lpdf(y | mu .* w, sigma);
}
generated quantities {
  real s2 = square(sigma);
}
