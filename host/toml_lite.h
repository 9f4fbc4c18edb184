// Minimal TOML reader for QuetzalPIC input files (C++14, no dependencies).
//
// The reference parses its inputs with cpptoml (src/Config_reader.cpp:26), which is not available here.  The input
// files only use: `key = integer | float | true/false | "string" | 'string'`, `[table]`, `[[array-of-tables]]` and `#`
// comments (inputs/test/*.toml), which is what this reader supports.  As in cpptoml, a floating-point key accepts an
// integer literal (`lx = 20`), but an integer key does not accept a float.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace toml_lite {

struct Value {
    enum Kind { INTEGER, FLOAT, BOOLEAN, STRING } kind = INTEGER;
    std::int64_t i = 0;
    double d = 0;
    bool b = false;
    std::string s;
};

struct Table {
    std::map<std::string, Value> values;
    std::map<std::string, std::shared_ptr<Table>> tables;
    std::map<std::string, std::vector<std::shared_ptr<Table>>> table_arrays;

    bool has_value(const std::string & k) const { return values.count(k) != 0; }
    bool contains(const std::string & k) const { return values.count(k) || tables.count(k) || table_arrays.count(k); }

    bool get(const std::string & k, double & out) const {
        auto it = values.find(k);
        if (it == values.end()) return false;
        if (it->second.kind == Value::FLOAT) out = it->second.d;
        else if (it->second.kind == Value::INTEGER) out = static_cast<double>(it->second.i);
        else return false;
        return true;
    }
    bool get(const std::string & k, int & out) const {
        auto it = values.find(k);
        if (it == values.end() || it->second.kind != Value::INTEGER) return false;
        out = static_cast<int>(it->second.i);
        return true;
    }
    bool get(const std::string & k, bool & out) const {
        auto it = values.find(k);
        if (it == values.end() || it->second.kind != Value::BOOLEAN) return false;
        out = it->second.b;
        return true;
    }
    bool get(const std::string & k, std::string & out) const {
        auto it = values.find(k);
        if (it == values.end() || it->second.kind != Value::STRING) return false;
        out = it->second.s;
        return true;
    }
};

inline std::string trim(const std::string & s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r' || s[a] == '\n')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) b--;
    return s.substr(a, b - a);
}

inline std::string strip_comment(const std::string & s) {
    char quote = 0;
    for (size_t i = 0; i < s.size(); i++) {
        if (quote) {
            if (s[i] == quote) quote = 0;
        } else if (s[i] == '"' || s[i] == '\'') {
            quote = s[i];
        } else if (s[i] == '#') {
            return s.substr(0, i);
        }
    }
    return s;
}

inline Value parse_value(const std::string & raw, int line) {
    Value v;
    const std::string s = trim(raw);
    if (s.empty()) throw std::runtime_error("TOML: empty value at line " + std::to_string(line));
    if ((s.front() == '"' || s.front() == '\'') && s.size() >= 2 && s.back() == s.front()) {
        v.kind = Value::STRING;
        v.s = s.substr(1, s.size() - 2);
        return v;
    }
    if (s == "true" || s == "false") {
        v.kind = Value::BOOLEAN;
        v.b = (s == "true");
        return v;
    }
    std::string digits;
    for (char c : s) {
        if (c != '_') digits.push_back(c);
    }
    const bool is_float = digits.find_first_of(".eE") != std::string::npos || digits.find("inf") != std::string::npos ||
                          digits.find("nan") != std::string::npos;
    char * end = nullptr;
    if (is_float) {
        v.kind = Value::FLOAT;
        v.d = std::strtod(digits.c_str(), &end);
    } else {
        v.kind = Value::INTEGER;
        v.i = std::strtoll(digits.c_str(), &end, 10);
    }
    if (end == digits.c_str() || *end != '\0') throw std::runtime_error("TOML: cannot parse '" + s + "' at line " + std::to_string(line));
    return v;
}

inline std::shared_ptr<Table> parse_file(const std::string & filename) {
    std::ifstream in(filename);
    if (!in.is_open()) throw std::runtime_error(filename + " could not be opened for parsing");
    auto root = std::make_shared<Table>();
    Table * current = root.get();
    std::string line;
    int line_no = 0;
    while (std::getline(in, line)) {
        line_no++;
        const std::string s = trim(strip_comment(line));
        if (s.empty()) continue;
        if (s.size() >= 4 && s.compare(0, 2, "[[") == 0 && s.compare(s.size() - 2, 2, "]]") == 0) {
            const std::string name = trim(s.substr(2, s.size() - 4));
            if (root->tables.count(name) || root->values.count(name)) throw std::runtime_error("TOML: key " + name + " redefined");
            auto t = std::make_shared<Table>();
            root->table_arrays[name].push_back(t);
            current = t.get();
        } else if (s.front() == '[' && s.back() == ']') {
            const std::string name = trim(s.substr(1, s.size() - 2));
            if (root->contains(name)) throw std::runtime_error("TOML: table " + name + " redefined");
            auto t = std::make_shared<Table>();
            root->tables[name] = t;
            current = t.get();
        } else {
            const size_t eq = s.find('=');
            if (eq == std::string::npos) throw std::runtime_error("TOML: expected key = value at line " + std::to_string(line_no));
            current->values[trim(s.substr(0, eq))] = parse_value(s.substr(eq + 1), line_no);
        }
    }
    return root;
}

}  // namespace toml_lite
