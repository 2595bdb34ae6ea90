/*
 * Shim of boost/numeric/odeint.hpp for compiling src/util/CPGEquations.cpp in place (oracle/ref): forwards to the
 * restatement of Boost 1.53's integrate() in oracle/tg_odeint.h (Boost is not vendored with the reference).
 */
#ifndef REF_SHIM_BOOST_NUMERIC_ODEINT_HPP
#define REF_SHIM_BOOST_NUMERIC_ODEINT_HPP
#include <iostream> /* Boost.Odeint pulls <iostream> in; CPGEquations.cpp relies on it for std::cout */
#include "../../../../tg_odeint.h"
namespace boost { namespace numeric { namespace odeint {
using tg_odeint::integrate;
}}}
#endif
