/* Test fixture: a grab bag of the supported densities and functions, with every kind of parameter bound,
   local arrays, compound assignments and a loop-carried local. */
data {
  int<lower=1> N;
  vector[N] t;
  vector<lower=0>[N] y;
  array[N] int<lower=0> k;
  real<lower=0, upper=1> phi;
}
parameters {
  real<lower=0> a;            // lognormal prior
  real<upper=3> b;            // upper bound only
  real<lower=0, upper=1> p;   // beta prior
  real c;
  vector<lower=0.5>[2] s;     // lower bound != 0
}
model {
  vector[N] m;
  real acc;
  a ~ lognormal(0.2, 0.7);
  target += normal_lpdf(b | 1, 2) + double_exponential_lpdf(c | 0, 1.5);
  p ~ beta(2.5, 1.5);
  s ~ gamma(3, 2);
  acc = 0;
  for (n in 1:N) {
    m[n] = a * exp(-square(t[n]) / s[1]) + pow(s[2], 1.5) * p;
    acc += m[n] * 0.1;                      // loop-carried
    acc *= 0.9;
    target += phi * (lognormal_lpdf(y[n] | log(m[n]) + 0.01 * acc, 0.3 + inv_logit(c))
                     + poisson_log_lpmf(k[n] | log_sum_exp(b, c) - 2));
  }
  target += -0.5 * acc * acc / 100;
}
