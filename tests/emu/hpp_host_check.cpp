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
        } else if (kind == "R") {
            // R <case>: result records of a fixed set of (data, err) values, hex
            static gofr::App app;
            static auto& item = app.Struct("main.Item").String("SKU", "sku").Int32("Qty", "qty").Int64("Big", "big").Bool("Ok", "ok").Int("N", "n").String("Note", "note", true);
            const int c = atoi(a.c_str());
            gofr::Result res;
            const gofr::App::StructType* ty = &item;
            if (c == 0) res = gofr::Result(std::string("Hello World!"));
            else if (c == 1) res = gofr::Result(gofr::Error{"db: connection refused"});
            else if (c == 2) res = gofr::Result();
            else if (c == 3) res = gofr::Result(gofr::ErrMissingFile());
            else if (c == 4) res = gofr::Result(gofr::Data(item({std::string("A-1"), int64_t(-3), int64_t(-5000000000LL), true, int64_t(1) << 40, std::string("fragile")})));
            else if (c == 5) res = gofr::Result(gofr::Data(item({std::string(""), int64_t(0), int64_t(0), false, int64_t(0), std::string("")})), gofr::Error{"partial"});
            else if (c == 6) { res = gofr::Result(gofr::Data(item({std::string("x"), int64_t(1)}))); }                         // wrong field count
            else if (c == 7) { res = gofr::Result(gofr::Data(item({int64_t(1), int64_t(1), int64_t(1), true, int64_t(1), std::string("")}))); }  // wrong kind
            else if (c == 8) { res = gofr::Result(gofr::Data(item({std::string("A"), int64_t(1), int64_t(1), true, int64_t(1), std::string("")}))); ty = nullptr; }  // route without a type
            else if (c == 9) res = gofr::Result(gofr::Data(std::string("s")), gofr::Error{"e"});                                 // string next to an error
            const std::string rec = app.ResultRecord(res, ty);
            for (unsigned char ch : rec) printf("%02x", ch);
            printf("\n");
        } else if (kind == "V") {
            // V <case>: rows of the wider data model (nested struct, pointer, slice, map, float64, bare slice), hex
            using gofr::List; using gofr::Map; using gofr::Nil; using gofr::Value;
            static gofr::App app;
            static auto& addr = app.Struct("main.Addr").String("City", "city").Int32("Zip", "zip", true).Float64("Geo", "geo").Slice();
            static auto& user = app.Struct("main.User").String("Name", "name").Float64("Score", "score").Struct("Home", addr, "home")
                                    .Struct("Work", addr, "work", true).Ptr().String("Tags", "tags").Slice().String("Attrs", "attrs", true).MapOf()
                                    .Struct("Hist", addr, "hist").Slice().Int64("N", "n", true).Ptr().Int64("Counts", "counts").MapOf();
            static auto& addrs = app.Bare("[]main.Addr").Struct("", addr).Slice();
            static auto& blob = app.Struct("main.Blob").Uint64("ID", "id").Bytes("Data", "data").Bytes("Sum", "sum", true).Float32("Ratio", "ratio")
                                    .Bytes("Parts", "parts").Slice().Uint64("UM", "um").MapOf().Float32("PF", "pf").Ptr().TimeField("At", "at").TimeField("Seen", "seen").Slice()
                                    .Struct("Kids", addr, "kids").SliceOfPtr().Int64("PI", "pi").SliceOfPtr();
            const int c = atoi(a.c_str());
            gofr::Result res;
            const gofr::App::StructType* ty = &user;
            if (c == 0) res = gofr::Result(gofr::Data(user({"bo<b>", 1.5e-7, addr({"Paris", 0, List{1.0, 2.5}}), Nil{}, List{"a", "b\n"},
                                                            Map{{{"z", "1"}, {"a", "2"}, {"aa", "3"}}}, List{addr({"X", 7, Nil{}}), addr({"Y", 0, List{}})},
                                                            int64_t(5), Nil{}})));
            else if (c == 1) res = gofr::Result(gofr::Data(user({"", 0.0, addr({"", 0, Nil{}}), addr({"W", 1, List{-0.5}}), Nil{}, Nil{}, Nil{}, Nil{},
                                                                 Map{{{"k", int64_t(-1)}}}})));
            else if (c == 2) { ty = &addrs; res = gofr::Result(gofr::Data(addrs({List{addr({"X", 7, Nil{}}), addr({"Y", 2, List{3.25}})}}))); }
            else if (c == 3) { ty = &addrs; res = gofr::Result(gofr::Data(addrs({Nil{}}))); }
            else if (c == 4) res = gofr::Result(gofr::Data(user({"n", 1.0, addr({"P", 0, Nil{}}), "not a struct", Nil{}, Nil{}, Nil{}, Nil{}, Nil{}})));  // wrong kind
            else if (c == 5) { ty = &blob; res = gofr::Result(gofr::Data(blob({uint64_t(18446744073709551615ull), std::string("\x00\xff\x10", 3), Nil{}, 0.1,
                                                                               List{std::string("ab"), Nil{}, std::string("")}, Map{{{"k", uint64_t(1) << 63}, {"j", int64_t(7)}}}, 2.5,
                                                                               gofr::Time{1709210096, 123456789, 19800}, List{gofr::Time{}, gofr::Time{0, 5, -3600}},
                                                                               List{Nil{}, addr({"K", 9, List{0.5}}), Nil{}}, List{int64_t(4), Nil{}}}))); }
            const std::string rec = app.ResultRecord(res, ty);
            for (unsigned char ch : rec) printf("%02x", ch);
            printf("\n");
        } else if (kind == "P") {
            std::string path, query;
            bool force = false;
            if (!gofr::App::ParseTarget(unhex(a), &path, &query, &force)) { printf("ERR\n"); continue; }
            printf("-");
            for (unsigned char ch : path) printf("%02x", ch);
            printf(" -");
            for (unsigned char ch : query) printf("%02x", ch);
            printf(" %d\n", force ? 1 : 0);
        } else if (kind == "T") {
            const auto names = gofr::detail::template_vars(unhex(a));
            for (size_t i = 0; i < names.size(); i++) printf("%s%s", i ? "," : "", names[i].c_str());
            printf("\n");
        }
    }
    return 0;
}
