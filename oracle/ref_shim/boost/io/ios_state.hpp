// Stand-in for <boost/io/ios_state.hpp> (test/test_utils.hpp:9,53): restores stream formatting on scope exit.
#pragma once
#include <ios>
namespace boost { namespace io {
class ios_all_saver {
  std::ios_base& s_; std::ios_base::fmtflags f_; std::streamsize p_, w_;
 public:
  explicit ios_all_saver(std::ios_base& s) : s_(s), f_(s.flags()), p_(s.precision()), w_(s.width()) {}
  ~ios_all_saver() { s_.flags(f_); s_.precision(p_); s_.width(w_); }
};
}}
