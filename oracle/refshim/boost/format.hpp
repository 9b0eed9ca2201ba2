// Stand-in for <boost/format.hpp> (TEST INFRASTRUCTURE, see oracle/refshim/README.md): "%s" / "%d" / "%04d"-style
// directives filled in order, .str().
#ifndef REFSHIM_BOOST_FORMAT
#define REFSHIM_BOOST_FORMAT
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
public:
    explicit format(const std::string &f) : f_(f) {}
    template <class T> format &operator%(const T &v) { std::ostringstream o; o << v; args_.push_back(o.str()); return *this; }
    std::string str() const
    {
        std::string out;
        size_t a = 0;
        for (size_t i = 0; i < f_.size(); ++i) {
            if (f_[i] != '%') { out.push_back(f_[i]); continue; }
            if (i + 1 < f_.size() && f_[i + 1] == '%') { out.push_back('%'); ++i; continue; }
            size_t j = i + 1;
            while (j < f_.size() && std::string("sdif").find(f_[j]) == std::string::npos) ++j;      // skip flags / width
            if (a < args_.size()) out += args_[a++];
            i = j;
        }
        return out;
    }
private:
    std::string f_;
    std::vector<std::string> args_;
};
}  // namespace boost
#endif
