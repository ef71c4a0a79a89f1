// TEST INFRASTRUCTURE ONLY: exposes the host-side helpers of include/gofr_b200.hpp (query parsing, template variable
// names, target parsing) on stdin/stdout so that tests/test_cpp_app.py can compare them with the oracle.
//   lines in:  Q <hex raw query> <hex key>   |  T <hex pattern>
//   lines out: hex value                     |  names joined by ','
#include <cstdio>
#include <iostream>
#include <sstream>

#include "../../include/gofr_b200.hpp"

static std::string unhex(const std::string& h) {
    std::string s;
    for (size_t i = 0; i + 1 < h.size(); i += 2) s.push_back((char)(gofr::detail::hexval((unsigned char)h[i]) * 16 + gofr::detail::hexval((unsigned char)h[i + 1])));
    return s;
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string kind, a, b;
        is >> kind >> a >> b;
        if (a == "-") a.clear();
        if (b == "-") b.clear();
        if (kind == "Q") {
            const std::string v = gofr::detail::query_get(unhex(a), unhex(b));
            for (unsigned char c : v) printf("%02x", c);
            printf("\n");
        } else if (kind == "T") {
            const auto names = gofr::detail::template_vars(unhex(a));
            for (size_t i = 0; i < names.size(); i++) printf("%s%s", i ? "," : "", names[i].c_str());
            printf("\n");
        }
    }
    return 0;
}
