// Stand-in for <boost/algorithm/string.hpp> (TEST INFRASTRUCTURE, see oracle/refshim/README.md): split + is_any_of.
#ifndef REFSHIM_BOOST_STRING
#define REFSHIM_BOOST_STRING
#include <string>
#include <vector>
namespace boost {
struct is_any_of_t { std::string set; };
inline is_any_of_t is_any_of(const std::string &s) { return is_any_of_t{s}; }
template <class Seq>
inline void split(std::vector<std::string> &out, const Seq &input, const is_any_of_t &pred)
{
    const std::string in(input);
    out.clear();
    std::string cur;
    for (char ch : in) {
        if (pred.set.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(ch);
    }
    out.push_back(cur);
}
}  // namespace boost
#endif
