// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the third-party `cpptoml` header, just large enough to let the UNMODIFIED
// reference sources under /root/reference/src compile into oracle/_ref/ (see oracle/Makefile).
// cpptoml itself is not installed in this image.  API surface = exactly what the reference calls
// (src/Config_reader.cpp:26-31,83,99,110-112,190-201,221-223,237-239; src/Config_reader.h:6,17).
//
// Supported TOML subset: `key = integer | float | true/false | "string" | 'string'`, `[table]`,
// `[[array-of-tables]]`, `#` comments.  get_as<double> accepts TOML integers (cpptoml does the same;
// the shipped inputs rely on it, e.g. inputs/test/laser_test.toml:1 `lx = 20`).
#pragma once

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace cpptoml {

template <class T>
class option {
public:
    option() : has_(false), value_() {}
    option(T v) : has_(true), value_(std::move(v)) {}
    explicit operator bool() const { return has_; }
    const T & operator*() const { return value_; }
    const T * operator->() const { return &value_; }

private:
    bool has_;
    T value_;
};

class parse_exception : public std::runtime_error {
public:
    explicit parse_exception(const std::string & msg) : std::runtime_error(msg) {}
};

class base {
public:
    virtual ~base() = default;
    virtual bool is_table() const { return false; }
    virtual bool is_table_array() const { return false; }
    virtual bool is_value() const { return false; }
};

enum class value_kind { integer, floating, boolean, string };

class scalar : public base {
public:
    value_kind kind = value_kind::integer;
    int64_t i = 0;
    double d = 0;
    bool b = false;
    std::string s;
    bool is_value() const override { return true; }
};

class table;

class table_array : public base {
public:
    bool is_table_array() const override { return true; }
    std::vector<std::shared_ptr<table>> & get() { return tables_; }
    const std::vector<std::shared_ptr<table>> & get() const { return tables_; }

private:
    std::vector<std::shared_ptr<table>> tables_;
};

class table : public base {
public:
    bool is_table() const override { return true; }

    bool contains(const std::string & key) const { return map_.count(key) != 0; }

    std::shared_ptr<base> get(const std::string & key) const {
        auto it = map_.find(key);
        if (it == map_.end()) throw std::out_of_range(key + " is not a valid key");
        return it->second;
    }

    std::shared_ptr<table> get_table(const std::string & key) const {
        auto it = map_.find(key);
        if (it == map_.end()) return nullptr;
        return std::dynamic_pointer_cast<table>(it->second);
    }

    std::shared_ptr<table_array> get_table_array(const std::string & key) const {
        auto it = map_.find(key);
        if (it == map_.end()) return nullptr;
        return std::dynamic_pointer_cast<table_array>(it->second);
    }

    template <class T>
    option<T> get_as(const std::string & key) const {
        auto it = map_.find(key);
        if (it == map_.end()) return {};
        auto v = std::dynamic_pointer_cast<scalar>(it->second);
        if (!v) return {};
        return convert<T>(*v);
    }

    void insert(const std::string & key, std::shared_ptr<base> v) { map_[key] = std::move(v); }

private:
    template <class T>
    static typename std::enable_if<std::is_same<T, bool>::value, option<T>>::type convert(const scalar & v) {
        if (v.kind == value_kind::boolean) return option<T>(v.b);
        return {};
    }
    template <class T>
    static typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value, option<T>>::type
    convert(const scalar & v) {
        if (v.kind == value_kind::integer) return option<T>(static_cast<T>(v.i));
        return {};
    }
    template <class T>
    static typename std::enable_if<std::is_floating_point<T>::value, option<T>>::type convert(const scalar & v) {
        if (v.kind == value_kind::floating) return option<T>(static_cast<T>(v.d));
        if (v.kind == value_kind::integer) return option<T>(static_cast<T>(v.i));
        return {};
    }
    template <class T>
    static typename std::enable_if<std::is_same<T, std::string>::value, option<T>>::type convert(const scalar & v) {
        if (v.kind == value_kind::string) return option<T>(v.s);
        return {};
    }

    std::map<std::string, std::shared_ptr<base>> map_;
};

namespace detail {

inline std::string trim(const std::string & s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
    return s.substr(a, b - a);
}

// strips a trailing comment that is not inside a quoted string
inline std::string strip_comment(const std::string & s) {
    char quote = 0;
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[i];
        if (quote) {
            if (c == quote) quote = 0;
        } else if (c == '"' || c == '\'') {
            quote = c;
        } else if (c == '#') {
            return s.substr(0, i);
        }
    }
    return s;
}

inline std::shared_ptr<scalar> parse_scalar(const std::string & raw, int line_no) {
    auto v = std::make_shared<scalar>();
    const std::string s = trim(raw);
    if (s.empty()) throw parse_exception("empty value at line " + std::to_string(line_no));
    if ((s.front() == '"' || s.front() == '\'') && s.size() >= 2 && s.back() == s.front()) {
        v->kind = value_kind::string;
        v->s = s.substr(1, s.size() - 2);
        return v;
    }
    if (s == "true" || s == "false") {
        v->kind = value_kind::boolean;
        v->b = (s == "true");
        return v;
    }
    std::string digits;
    for (char c : s) {
        if (c != '_') digits.push_back(c);
    }
    const bool is_float = digits.find_first_of(".eE") != std::string::npos || digits == "inf" || digits == "nan" ||
                          digits == "+inf" || digits == "-inf";
    char * end = nullptr;
    if (is_float) {
        v->kind = value_kind::floating;
        v->d = std::strtod(digits.c_str(), &end);
    } else {
        v->kind = value_kind::integer;
        v->i = std::strtoll(digits.c_str(), &end, 10);
    }
    if (end == digits.c_str() || *end != '\0') {
        throw parse_exception("cannot parse value '" + s + "' at line " + std::to_string(line_no));
    }
    return v;
}

}  // namespace detail

inline std::shared_ptr<table> parse_file(const std::string & filename) {
    std::ifstream in(filename);
    if (!in.is_open()) throw parse_exception(filename + " could not be opened for parsing");

    auto root = std::make_shared<table>();
    std::shared_ptr<table> current = root;
    std::string line;
    int line_no = 0;
    while (std::getline(in, line)) {
        line_no++;
        std::string s = detail::trim(detail::strip_comment(line));
        if (s.empty()) continue;
        if (s.size() >= 4 && s.substr(0, 2) == "[[" && s.substr(s.size() - 2) == "]]") {
            const std::string name = detail::trim(s.substr(2, s.size() - 4));
            std::shared_ptr<table_array> arr;
            if (root->contains(name)) {
                arr = root->get_table_array(name);
                if (!arr) throw parse_exception("key " + name + " redefined at line " + std::to_string(line_no));
            } else {
                arr = std::make_shared<table_array>();
                root->insert(name, arr);
            }
            current = std::make_shared<table>();
            arr->get().push_back(current);
        } else if (s.front() == '[' && s.back() == ']') {
            const std::string name = detail::trim(s.substr(1, s.size() - 2));
            if (root->contains(name)) {
                throw parse_exception("table " + name + " redefined at line " + std::to_string(line_no));
            }
            current = std::make_shared<table>();
            root->insert(name, current);
        } else {
            const size_t eq = s.find('=');
            if (eq == std::string::npos) {
                throw parse_exception("expected key = value at line " + std::to_string(line_no));
            }
            const std::string key = detail::trim(s.substr(0, eq));
            current->insert(key, detail::parse_scalar(s.substr(eq + 1), line_no));
        }
    }
    return root;
}

}  // namespace cpptoml
