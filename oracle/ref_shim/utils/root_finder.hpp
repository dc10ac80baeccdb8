// oracle/ref_shim/utils/root_finder.hpp — shadows the reference's utils/root_finder.hpp (a polynomial root finder with a
// large Eigen surface) when `utils/trajectory.hpp` is included whole into the pin library.  Only Piece<D>/Trajectory<D>
// member templates that the hot path never instantiates (max-rate checks) name these functions; declarations suffice.
#pragma once
#include <set>
#include <Eigen/Eigen>
namespace RootFinder {
template <class... A> Eigen::VectorXd polySqr(A &&...);
template <class... A> double polyVal(A &&...);
template <class... A> std::set<double> solvePolynomial(A &&...);
template <class... A> int countRoots(A &&...);
}  // namespace RootFinder
