/*
 * Restatement of the one Boost.Odeint entry point the reference uses (TEST INFRASTRUCTURE, oracle/):
 *     integrate(system, x, t0, t1, dt, observer)                         src/util/CPGEquations.cpp:253
 * Boost 1.53 (conf/default/boost.conf.default) is not vendored with the reference and not in this image, so its
 * published algorithm is restated: integrate() = integrate_adaptive(controlled_runge_kutta<runge_kutta_dopri5<State>>())
 * with the default error checker (eps_abs = eps_rel = 1e-6, a_x = a_dxdt = 1):
 *   - Dormand-Prince 5(4), FSAL: the derivative at the start of a step is the last stage of the previous accepted step;
 *     a freshly constructed stepper evaluates it once ("initialize").
 *   - err_i = |xerr_i| / (eps_abs + eps_rel * (a_x |x_i| + a_dxdt dt |dxdt_i|)), max over i.
 *   - err > 1: reject, dt *= max(0.9 err^(-1/3), 0.2);  err < 0.5: accept, dt *= 0.9 max(5^-5, err)^(-1/5) (<= 5x);
 *     otherwise accept with dt unchanged.
 *   - loop while t1 - t > epsilon; dt is cut to land on t1; the observer sees the state before every step and at the end;
 *     1000 failed trials in a row throw std::overflow_error.
 */
#ifndef TG_ODEINT_H
#define TG_ODEINT_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <stdexcept>

namespace tg_odeint {

namespace detail {

template <class State>
struct dopri5_controlled {
    State k1, k2, k3, k4, k5, k6, k7, xtmp, xnew, xerr;
    bool first;
    dopri5_controlled() : first(true) {}

    void resize(const State& x) {
        const std::size_t n = x.size();
        k1.resize(n); k2.resize(n); k3.resize(n); k4.resize(n); k5.resize(n); k6.resize(n); k7.resize(n);
        xtmp.resize(n); xnew.resize(n); xerr.resize(n);
    }

    /* returns true on an accepted step */
    template <class System>
    bool try_step(System& sys, State& x, double& t, double& dt) {
        if (first) {
            resize(x);
            sys(x, k1, t);
            first = false;
        }
        const std::size_t n = x.size();
        const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
        const double b21 = 1.0 / 5.0;
        const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
        const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
        const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
        const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0, b65 = -5103.0 / 18656.0;
        const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
        const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
                     dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;

        for (std::size_t i = 0; i < n; ++i) xtmp[i] = 1.0 * x[i] + (dt * b21) * k1[i];
        sys(xtmp, k2, t + dt * a2);
        for (std::size_t i = 0; i < n; ++i) xtmp[i] = 1.0 * x[i] + (dt * b31) * k1[i] + (dt * b32) * k2[i];
        sys(xtmp, k3, t + dt * a3);
        for (std::size_t i = 0; i < n; ++i) xtmp[i] = 1.0 * x[i] + (dt * b41) * k1[i] + (dt * b42) * k2[i] + (dt * b43) * k3[i];
        sys(xtmp, k4, t + dt * a4);
        for (std::size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (dt * b51) * k1[i] + (dt * b52) * k2[i] + (dt * b53) * k3[i] + (dt * b54) * k4[i];
        sys(xtmp, k5, t + dt * a5);
        for (std::size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (dt * b61) * k1[i] + (dt * b62) * k2[i] + (dt * b63) * k3[i] + (dt * b64) * k4[i] + (dt * b65) * k5[i];
        sys(xtmp, k6, t + dt);
        for (std::size_t i = 0; i < n; ++i)
            xnew[i] = 1.0 * x[i] + (dt * c1) * k1[i] + (dt * c3) * k3[i] + (dt * c4) * k4[i] + (dt * c5) * k5[i] + (dt * c6) * k6[i];
        sys(xnew, k7, t + dt);
        for (std::size_t i = 0; i < n; ++i)
            xerr[i] = (dt * dc1) * k1[i] + (dt * dc3) * k3[i] + (dt * dc4) * k4[i] + (dt * dc5) * k5[i] + (dt * dc6) * k6[i] + (dt * dc7) * k7[i];

        const double eps_abs = 1e-6, eps_rel = 1e-6, a_x = 1.0, a_dxdt = 1.0;
        double max_rel_err = 0.0;
        for (std::size_t i = 0; i < n; ++i) {
            const double e = std::fabs(xerr[i]) / (eps_abs + eps_rel * (a_x * std::fabs(x[i]) + (a_dxdt * dt) * std::fabs(k1[i])));
            max_rel_err = std::max(max_rel_err, e);
        }
        if (max_rel_err > 1.0) {
            dt *= std::max(9.0 / 10.0 * std::pow(max_rel_err, -1.0 / (4.0 - 1.0)), 1.0 / 5.0);
            return false;
        }
        if (max_rel_err < 0.5) {
            max_rel_err = std::max(std::pow(5.0, -5.0), max_rel_err);
            t += dt;
            dt *= 9.0 / 10.0 * std::pow(max_rel_err, -1.0 / 5.0);
        } else {
            t += dt;
        }
        for (std::size_t i = 0; i < n; ++i) { x[i] = xnew[i]; k1[i] = k7[i]; }
        return true;
    }
};

} // namespace detail

template <class System, class State, class Time, class Observer>
std::size_t integrate(System system, State& start_state, Time start_time, Time end_time, Time dt, Observer observer) {
    detail::dopri5_controlled<State> stepper;
    const std::size_t max_attempts = 1000;
    std::size_t count = 0;
    double t = start_time;
    while (end_time - t > std::numeric_limits<double>::epsilon()) {
        observer(start_state, t);
        if ((t + dt) - end_time > std::numeric_limits<double>::epsilon()) dt = end_time - t;
        std::size_t trials = 0;
        bool ok;
        do {
            ok = stepper.try_step(system, start_state, t, dt);
            ++trials;
        } while (!ok && trials < max_attempts);
        if (trials == max_attempts) throw std::overflow_error("Integrate adaptive : Maximal number of iterations reached. A step size could not be found.");
        ++count;
    }
    observer(start_state, t);
    return count;
}

} // namespace tg_odeint
#endif
