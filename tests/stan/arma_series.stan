// Test fixture (written for this repo): the ARMA(1,1) density of the built-in "arma" device model phrased with whole-series
// vector locals and ONE vectorised likelihood statement after the loop -- the style the generator fuses back into the loop.
data {
  int<lower=2> T;
  vector[T] y;
  real<lower=0, upper=1> phi;
}
parameters {
  real mu;
  real beta;
  real theta;
  real<lower=0> sigma;
}
model {
  vector[T] pred;
  vector[T] resid;
  target += normal_lpdf(mu | 0, 10) + normal_lpdf(beta | 0, 2) + normal_lpdf(theta | 0, 2) + cauchy_lpdf(sigma | 0, 2.5);
  pred[1] = mu + beta * mu;
  resid[1] = y[1] - pred[1];
  for (t in 2:T) {
    pred[t] = mu + beta * y[t - 1] + theta * resid[t - 1];
    resid[t] = y[t] - pred[t];
  }
  target += phi * normal_lpdf(resid | 0, sigma);
}
