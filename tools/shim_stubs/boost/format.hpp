#pragma once
#include <sstream>
#include <string>
namespace boost {
// accepts the same expressions as boost::format; the formatting itself is irrelevant to a syntax check
class format {
public:
    explicit format(const char* s) : _s(s ? s : "") {}
    explicit format(const std::string& s) : _s(s) {}
    template <typename T> format& operator%(const T& v) { std::ostringstream ss; ss << v; _s += " " + ss.str(); return *this; }
    std::string str() const { return _s; }
private:
    std::string _s;
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
}
