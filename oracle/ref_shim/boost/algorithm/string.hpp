// Stand-in for <boost/algorithm/string.hpp>: Utils.hpp:10 includes it and uses nothing; UserInput.hpp:16
// calls boost::split(parts, s, boost::is_any_of(delims)).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of { std::string set; explicit is_any_of(const std::string& s) : set(s) {} };
inline void split(std::vector<std::string>& out, const std::string& s, const is_any_of& d) {
  out.clear();
  std::string cur;
  for (char ch : s) {
    if (d.set.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(ch);
  }
  out.push_back(cur);
}
}
