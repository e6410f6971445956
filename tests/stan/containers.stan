/* Test fixture: container-valued expressions -- row_vector * vector, elementwise .* and ./, sum / mean / dot_product /
   dot_self, unary functions of vectors, densities of vector expressions, if / else on data and on parameters. */
data {
  int<lower=1> N;
  int<lower=1> K;
  array[N] row_vector[K] x;
  array[N] int<lower=0, upper=1> z;
  array[N] int<lower=0> trials;
  array[N] int<lower=0> wins;
  vector[N] t;
  int<lower=0, upper=1> robust;
  real<lower=0, upper=1> phi;
}
parameters {
  vector[K] b;
  real<lower=0> tau;
  real c;
  real<lower=0, upper=1> p;
}
model {
  vector[N] eta;
  b ~ normal(0, 2);
  tau ~ weibull(1.5, 2);
  c ~ logistic(0.5, 1.5);
  wins ~ binomial(trials, p);
  for (n in 1:N)
    eta[n] = x[n] * b + c;
  if (robust == 1) {
    target += phi * student_t_lpdf(t | 4, eta .* t - mean(eta), tau);
  } else {
    target += phi * normal_lpdf(t | eta .* t - mean(eta), tau);
  }
  z ~ bernoulli_logit(eta ./ (1 + tau));
  target += -0.5 * dot_self(b) / 10 + 0.01 * sum(exp(-eta)) - 0.1 * dot_product(b, b .* b);
  if (c > 0 && !(tau >= 10))
    target += -c;
  else
    target += c;
}
