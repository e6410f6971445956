// Test fixture (written for this repo): a tempered ARMA(1,1) likelihood with the same priors and recurrence as the
// built-in "arma" device model, phrased with a rolling residual instead of whole-series vectors.
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
  real resid;
  real pred;
  mu ~ normal(0, 10);
  target += -0.5 * log(2 * 3.141592653589793) - log(10.0);   // `~` dropped this constant; the built-in model keeps it
  target += normal_lpdf(beta | 0, 2) + normal_lpdf(theta | 0, 2);
  target += cauchy_lpdf(sigma | 0, 2.5);
  resid = y[1] - (mu + beta * mu);
  target += phi * normal_lpdf(resid | 0, sigma);
  for (t in 2:T) {
    pred = mu + beta * y[t - 1] + theta * resid;
    resid = y[t] - pred;
    target += phi * normal_lpdf(resid | 0, sigma);
  }
}
