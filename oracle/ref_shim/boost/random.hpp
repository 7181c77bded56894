// boost/random.hpp — the two Boost.Random names include/havoqgt/rmat_edge_generator.hpp uses (Boost 1.57 is the version the
// reference's build scripts pin, README.md:6; it is not vendored and not on this image).  Restated from the published
// algorithms, nothing else:
//   boost::mt19937              mersenne_twister_engine<uint32_t,32,624,397,31,0x9908b0df,11,0xffffffff,7,0x9d2c5680,15,
//                               0xefc60000,18,1812433253> — the same parameter set as std::mt19937 (both are "MT19937" of
//                               Matsumoto & Nishimura; the 10000th output of a default-seeded engine is 4123659995 in both)
//   boost::uniform_01<Engine>   the backward-compatible form taken when the template argument is an ENGINE, not a real
//                               type (boost/random/uniform_01.hpp, detail::backward_compatible_uniform_01): it keeps a COPY
//                               of the engine and returns  double(x - min) * (1 / (double(max - min) + 1)) = x * 2^-32,
//                               drawing again in the (here impossible) case that the product rounds to 1.
#pragma once
#include <cstdint>
#include <random>

namespace boost {

typedef std::mt19937 mt19937;

template <class Engine, class RealType = double>
class uniform_01 {
 public:
  typedef RealType result_type;
  explicit uniform_01(Engine rng) : _rng(rng), _factor(RealType(1) / (RealType((_rng.max)() - (_rng.min)()) + RealType(1))) {}
  result_type operator()() {
    for (;;) {
      const result_type r = result_type(_rng() - (_rng.min)()) * _factor;
      if (r < result_type(1)) return r;
    }
  }

 private:
  Engine _rng;  // by value: the generator of rmat_edge_generator.hpp copies its seeded engine into every iterator
  RealType _factor;
};

}  // namespace boost
