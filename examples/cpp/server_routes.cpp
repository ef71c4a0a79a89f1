// The reference's own route test, written against the C++ stand-in of the app API (include/gofr_b200.hpp):
// pkg/gofr/gofr_test.go:40-105 (TestGofr_ServerRoutes) plus the handlers of examples/http-server/main.go:29-43.
// Prints one line per case — "<status>\t<body>" — which tests/test_cpp_app.py compares with the oracle; exits non-zero
// when a response is not what the reference's test expects.
//   g++ -std=c++17 -Iinclude examples/cpp/server_routes.cpp gofr_b200/libgofr_b200.so -o server_routes
#include <cmath>
#include <cstdio>
#include <string>

#include "gofr_b200.hpp"

static std::string body_of(const std::string& wire) {
    const size_t p = wire.find("\r\n\r\n");
    return p == std::string::npos ? wire : wire.substr(p + 4);
}

int main() {
    const std::string helloWorld = "Hello World!";
    gofr::App g;  // g := New()

    g.GET("/hello", [&](gofr::Context&) -> gofr::Result { return helloWorld; });
    g.add("GET", "/hello2", [&](gofr::Context&) -> gofr::Result { return helloWorld; });  // using add() func
    g.PUT("/hello", [&](gofr::Context&) -> gofr::Result { return helloWorld; });
    g.POST("/hello", [&](gofr::Context&) -> gofr::Result { return helloWorld; });
    g.GET("/params", [](gofr::Context& c) -> gofr::Result { return "Hello " + c.Param("name") + "!"; });
    g.DELETE("/delete", [](gofr::Context&) -> gofr::Result { return "Success"; });
    // examples/http-server/main.go:29-43
    g.GET("/greet", [](gofr::Context& c) -> gofr::Result {
        std::string name = c.Param("name");
        if (name.empty()) name = "World";
        return "Hello " + name + "!";
    });
    g.GET("/error", [](gofr::Context&) -> gofr::Result { return gofr::Error{"some error occurred"}; });
    // beyond the reference's test: path variables, a struct, nil, a missing file, a panic
    auto& person = g.Struct("main.Person").Int("ID", "id").String("Name", "name").Bool("Admin", "admin", true);
    g.GET("/users/{id:[0-9]+}/posts/{slug}", [](gofr::Context& c) -> gofr::Result {
        return "user " + c.PathParam("id") + " post " + c.PathParam("slug");
    });
    g.GET("/person/{name}", [&](gofr::Context& c) -> gofr::Result {
        const std::string n = c.PathParam("name");
        if (n == "nobody") return gofr::Result(person({int64_t(0), n, false}), gofr::Error{"partial <result>"});
        return gofr::Data(person({int64_t(n.size()), n, n == "root"}));
    }, &person);
    g.GET("/nil", [](gofr::Context&) -> gofr::Result { return gofr::Result(); });
    g.GET("/file", [](gofr::Context&) -> gofr::Result { return gofr::ErrMissingFile(); });
    g.GET("/panic", [](gofr::Context&) -> gofr::Result { throw std::runtime_error("boom"); });
    g.POST("/echo", [](gofr::Context& c) -> gofr::Result { return c.Body(); });
    // Context.Bind (pkg/gofr/context.go:52-54; the shape of pkg/gofr/context_test.go:23-49): bind the body, return the struct
    g.POST("/people", [](gofr::Context& c) -> gofr::Result {
        gofr::StructValue p;
        if (auto err = c.Bind(&p)) return *err;
        return gofr::Data(p);
    }, &person);
    g.Binds(person);
    // response.Raw (pkg/gofr/http/response/raw.go:3-5): the data without the envelope; an error only picks the status
    g.GET("/raw", [](gofr::Context&) -> gofr::Result { return gofr::Raw("just <text>"); });
    g.GET("/rawnil", [](gofr::Context&) -> gofr::Result { return gofr::Raw(); });
    g.GET("/rawperson", [&](gofr::Context&) -> gofr::Result { return gofr::Raw(gofr::Data(person({int64_t(9), std::string("raw"), true}))); }, &person);
    g.GET("/rawerr", [](gofr::Context&) -> gofr::Result { return gofr::Result(gofr::Raw("x"), gofr::Error{"ignored"}); });
    // what encoding/json does with the rest of Go's data model (responder.go:32-40): float64, nested structs, pointers,
    // slices, map[string]T, a handler returning a slice rather than a struct; NaN makes Encode fail (no body)
    auto& addr = g.Struct("main.Addr").String("City", "city").Int32("Zip", "zip", true).Float64("Geo", "geo").Slice();
    auto& user = g.Struct("main.User").String("Name", "name").Float64("Score", "score").Struct("Home", addr, "home")
                     .Struct("Work", addr, "work", true).Ptr().String("Tags", "tags").Slice().String("Attrs", "attrs", true).MapOf();
    auto& addrs = g.Bare("[]main.Addr").Struct("", addr).Slice();
    g.GET("/user/{name}", [&](gofr::Context& c) -> gofr::Result {
        const std::string n = c.PathParam("name");
        if (n == "nan") return gofr::Data(user({n, std::nan(""), addr({"", 0, gofr::Nil{}}), gofr::Nil{}, gofr::Nil{}, gofr::Nil{}}));
        return gofr::Data(user({n, 1.5e-7, addr({"Paris", 75001, gofr::List{48.8566, 2.3522}}), addr({"Lyon", 0, gofr::List{}}),
                                gofr::List{"a", "b<c>"}, gofr::Map{{{"z", "1"}, {"a", "2"}}}}));
    }, &user);
    g.GET("/addrs", [&](gofr::Context& c) -> gofr::Result {
        if (c.Param("none") == "1") return gofr::Data(addrs({gofr::Nil{}}));
        return gofr::Data(addrs({gofr::List{addr({"X", 7, gofr::Nil{}}), addr({"Y", 0, gofr::List{1e21, -0.0}})}}));
    }, &addrs);

    g.Run(0);

    struct Case { const char* method; const char* target; const char* body; int status; const char* want; };
    const Case cases[] = {
        {"GET", "/hello", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"GET", "/hello2", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"PUT", "/hello", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"POST", "/hello", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"GET", "/params?name=Vikash", "", 200, "{\"data\":\"Hello Vikash!\"}\n"},
        {"DELETE", "/delete", "", 200, "{\"data\":\"Success\"}\n"},
        {"GET", "/greet", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"GET", "/greet?name=a%26b+c&name=second", "", 200, "{\"data\":\"Hello a\\u0026b c!\"}\n"},
        {"GET", "/greet?name=%zz&x=1", "", 200, "{\"data\":\"Hello World!\"}\n"},
        {"GET", "/error", "", 500, "{\"error\":{\"message\":\"some error occurred\"}}\n"},
        {"GET", "/users/42/posts/hello-world", "", 200, "{\"data\":\"user 42 post hello-world\"}\n"},
        {"GET", "/users/4x2/posts/p", "", 404, "{\"error\":{\"message\":\"http: no such file\"}}\n"},
        {"GET", "/person/root", "", 200, "{\"data\":{\"id\":4,\"name\":\"root\",\"admin\":true}}\n"},
        {"GET", "/person/al%20ice", "", 200, "{\"data\":{\"id\":6,\"name\":\"al ice\"}}\n"},
        {"GET", "/person/nobody", "", 500, "{\"error\":{\"message\":\"partial \\u003cresult\\u003e\"},\"data\":{\"id\":0,\"name\":\"nobody\"}}\n"},
        {"GET", "/nil", "", 200, "{}\n"},
        {"GET", "/file", "", 404, "{\"error\":{\"message\":\"http: no such file\"}}\n"},
        {"GET", "/panic", "", 500, "{\"code\":500,\"message\":\"Some unexpected error has occurred\",\"status\":\"ERROR\"}\n"},
        {"POST", "/echo", "line1\n\"quoted\" <tag>", 200, "{\"data\":\"line1\\n\\\"quoted\\\" \\u003ctag\\u003e\"}\n"},
        {"GET", "/hello/", "", 404, "{\"error\":{\"message\":\"http: no such file\"}}\n"},  // cleanPath keeps the slash; no StrictSlash
        {"GET", "/a/../hello", "", 301, nullptr},
        {"GET", "//hello", "", 301, nullptr},
        {"OPTIONS", "/hello", "", 200, ""},
        {"PATCH", "/hello", "", 404, "{\"error\":{\"message\":\"http: no such file\"}}\n"},
        {"GET", "/.well-known/health", "", 200, "{\"data\":{}}\n"},
        {"POST", "/people", "{\"id\":1,\"name\":\"Bob\"}", 200, "{\"data\":{\"id\":1,\"name\":\"Bob\"}}\n"},
        {"POST", "/people", "{\"ID\":7,\"NAME\":\"caf\\u00e9 \\\"q\\\"\",\"admin\":true,\"extra\":[1,{\"a\":2}]}", 200, "{\"data\":{\"id\":7,\"name\":\"caf\xc3\xa9 \\\"q\\\"\",\"admin\":true}}\n"},
        {"POST", "/people", "{\"id\":\"x\"}", 500, "{\"error\":{\"message\":\"json: cannot unmarshal string into Go struct field Person.id of type int\"}}\n"},
        {"POST", "/people", "{bad", 500, "{\"error\":{\"message\":\"invalid character 'b' looking for beginning of object key string\"}}\n"},
        {"POST", "/people", "", 500, "{\"error\":{\"message\":\"unexpected end of JSON input\"}}\n"},
        {"GET", "/raw", "", 200, "\"just \\u003ctext\\u003e\"\n"},
        {"GET", "/rawnil", "", 200, "null\n"},
        {"GET", "/rawperson", "", 200, "{\"id\":9,\"name\":\"raw\",\"admin\":true}\n"},
        {"GET", "/rawerr", "", 500, "\"x\"\n"},
        {"GET", "/user/bob", "", 200, "{\"data\":{\"name\":\"bob\",\"score\":1.5e-7,\"home\":{\"city\":\"Paris\",\"zip\":75001,\"geo\":[48.8566,2.3522]},"
                                      "\"work\":{\"city\":\"Lyon\",\"geo\":[]},\"tags\":[\"a\",\"b\\u003cc\\u003e\"],\"attrs\":{\"a\":\"2\",\"z\":\"1\"}}}\n"},
        {"GET", "/user/nan", "", 200, ""},
        {"GET", "/addrs", "", 200, "{\"data\":[{\"city\":\"X\",\"zip\":7,\"geo\":null},{\"city\":\"Y\",\"geo\":[1e+21,-0]}]}\n"},
        {"GET", "/addrs?none=1", "", 200, "{\"data\":null}\n"},
    };
    std::vector<gofr::App::Request> reqs;
    for (auto& c : cases) {
        gofr::App::Request r;
        r.method = c.method;
        r.target = c.target;
        r.body = c.body;
        std::array<uint8_t, 16> id{};
        for (int k = 0; k < 16; k++) id[k] = (uint8_t)(reqs.size() * 16 + k);
        r.trace_id = id;
        reqs.push_back(r);
    }
    const auto resp = g.Serve(reqs, 1700000000);
    int bad = 0;
    for (size_t i = 0; i < resp.size(); i++) {
        const std::string body = body_of(resp[i].bytes);
        const bool ok = resp[i].status == cases[i].status && (!cases[i].want || body == cases[i].want) &&
                        resp[i].bytes.rfind("HTTP/1.1 ", 0) == 0;
        if (!ok) { bad++; fprintf(stderr, "case %zu %s %s: status %d body %s\n", i, cases[i].method, cases[i].target, resp[i].status, body.c_str()); }
    }
    // one request at a time gives the same bytes as the batch
    for (size_t i = 0; i < reqs.size(); i += 5)
        if (g.ServeHTTP(reqs[i], 1700000000).bytes != resp[i].bytes) { bad++; fprintf(stderr, "case %zu: ServeHTTP differs from Serve\n", i); }
    // machine-readable dump for the oracle comparison: hex of every response
    for (size_t i = 0; i < resp.size(); i++) {
        printf("%d ", resp[i].status);
        for (unsigned char ch : resp[i].bytes) printf("%02x", ch);
        printf("\n");
    }
    return bad ? 1 : 0;
}
