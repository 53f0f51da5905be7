// Starting density of the liquid density iterations of the forward kernels.
//
// Every solve begins with a Newton iteration of a liquid density from the dense side (pure VLE, liquid density at
// (T, p), the liquid of a bubble / dew point).  Started at a fixed packing fraction it takes 4-5 of the 12-13
// Helmholtz evaluations of a state.  For non-associating, non-polar fluids the saturated liquid density is a
// universal function of the reduced temperature and the chain length:
//     rho_L / rho_max = F(tau, 1/m),   tau = (T / eps) / t_c(m),   rho_max = 0.5 / (pi/6 m sigma^3),
// fitted (oracle/fits/fit_liquid_start.py, oracle CPU solver, m in [1, 8], tau in [0.28, 0.97]; 15 terms in
// sqrt(1 - tau) and 1/m, error -4.0 % .. +1.9 %) and used, scaled by 1.02 towards the dense side, as the start:
// the iteration then needs 1-2 evaluations.  Mixtures use it with one-fluid (van der Waals) values of m and eps.
// It is a STARTING VALUE only: the iteration converges to the same root, states outside the fitted range,
// associating and polar states start where they did before, and a start that lands in the mechanically
// unstable range (dp/drho <= 0) is discarded for the old one.
#pragma once
#include "ad.cuh"

namespace feos {

// returns rho_L / rho_max, or 0 outside the fitted range
FD double liquid_start_fraction(double tstar, double m) {
  if (!(m >= 1.0) || !(m <= 8.0) || !(tstar > 0.0)) return 0.0;
  const double im = 1.0 / m;
  // critical temperature T_c / eps of the non-associating fluid, polynomial in 1/m (max error 0.3 %)
  double tc = 1.0556304946e+01;
  tc = tc * im - 4.0149642961e+01;
  tc = tc * im + 6.3360334370e+01;
  tc = tc * im - 5.4828173121e+01;
  tc = tc * im + 2.9769064610e+01;
  tc = tc * im - 1.1548226103e+01;
  tc = tc * im + 4.1160054607e+00;
  const double tau = tstar / tc;
  if (!(tau >= 0.28) || !(tau <= 0.97)) return 0.0;
  const double w = sqrt(1.0 - tau);
  // sum_a w^a (c_a0 + c_a1 / m + c_a2 / m^2), a = 0 .. 4
  const double c0 = 2.9411845664e-01 + im * (1.3057393127e-01 + im * -2.3164181408e-01);
  const double c1 = 1.7260923390e-01 + im * (8.1398119336e-01 + im * 8.7669948074e-01);
  const double c2 = 2.5513322118e+00 + im * (-2.8959835953e+00 + im * -2.2763772930e+00);
  const double c3 = -4.2172903482e+00 + im * (3.6181523797e+00 + im * 3.1945133566e+00);
  const double c4 = 2.4036878755e+00 + im * (-1.6780431026e+00 + im * -1.6749103578e+00);
  const double f = c0 + w * (c1 + w * (c2 + w * (c3 + w * c4)));
  return fmin(1.02 * f, 0.98);
}

}  // namespace feos
