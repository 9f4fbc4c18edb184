// Host-side input contract of QuetzalPIC, kept for the drop-in driver (C++14).
//
// Mirrors what the reference's Config_reader does with a TOML input (reference src/Config_reader.cpp:18-262): same keys,
// same defaults, same "name = value[ (default)]" echo, same built-in species table (:19-22), same profile functions
// (src/profiles.cpp:4-64) and unit conversions (src/Plasma_units.h, src/constants.h).  Written from the file format and
// the observable stdout, not from the reference's code structure.
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "toml_lite.h"

namespace qhost {

// ---- "{}" formatting of fmt for doubles: the shortest decimal string that round-trips
inline std::string fmt_double(double v) {
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    // shortest significand that round-trips
    char buf[64];
    int prec = 1;
    for (; prec <= 17; prec++) {
        std::snprintf(buf, sizeof(buf), "%.*e", prec - 1, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    std::string sci = buf;                       // [-]d[.ddd]e[+-]XX
    const size_t epos = sci.find('e');
    const int exp10 = std::atoi(sci.c_str() + epos + 1);
    std::string mant = sci.substr(0, epos);
    std::string sign;
    if (mant[0] == '-') {
        sign = "-";
        mant = mant.substr(1);
    }
    std::string digits;
    for (char c : mant) {
        if (c != '.') digits.push_back(c);
    }
    // fmt switches to exponent notation outside [1e-4, 1e16)
    if (exp10 < -4 || exp10 >= 16) {
        std::string out = sign + digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char e[16];
        std::snprintf(e, sizeof(e), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        return out + e;
    }
    if (exp10 >= static_cast<int>(digits.size()) - 1) return sign + digits + std::string(exp10 - (digits.size() - 1), '0');
    if (exp10 >= 0) return sign + digits.substr(0, exp10 + 1) + "." + digits.substr(exp10 + 1);
    return sign + "0." + std::string(-exp10 - 1, '0') + digits;
}
inline std::string fmt_value(double v) { return fmt_double(v); }
inline std::string fmt_value(int v) { return std::to_string(v); }
inline std::string fmt_value(bool v) { return v ? "true" : "false"; }
inline std::string fmt_value(const std::string & v) { return v; }
inline std::string fmt_g4(double v) {  // "{:.4g}"
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.4g", v);
    return buf;
}

inline std::string fmt_g3(double v) {  // "{:.3g}"
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.3g", v);
    return buf;
}

class Config_exception : public std::runtime_error {
public:
    explicit Config_exception(const std::string & m) : std::runtime_error(m) {}
};

// ---- constants and units (reference src/constants.h, src/Plasma_units.h)
constexpr double PI = 3.141592653589793238;
constexpr double ELECTRON_MASS_CGS = 9.1093837015e-28;
constexpr double PLANCK_CONST_CGS = 6.62607015e-27;
constexpr double PLANCK_CONST_BAR_CGS = PLANCK_CONST_CGS / (2 * PI);
constexpr double SPEED_OF_LIGHT_CGS = 29979245800.0;
constexpr double ELEMENTARY_CHARGE_CGS = 4.803204672997e-10;

struct Plasma_units {
    double frequency, plasma_wavenumber, wavelength, density, number_density_norm, field_schwinger;
    explicit Plasma_units(double base_frequency_SI = 1.0) {
        frequency = base_frequency_SI;
        plasma_wavenumber = frequency / SPEED_OF_LIGHT_CGS;
        wavelength = 2 * PI / plasma_wavenumber;
        density = ELECTRON_MASS_CGS * frequency * frequency / (4 * PI * ELEMENTARY_CHARGE_CGS * ELEMENTARY_CHARGE_CGS);
        number_density_norm = density / plasma_wavenumber / plasma_wavenumber / plasma_wavenumber;
        field_schwinger = ELECTRON_MASS_CGS * SPEED_OF_LIGHT_CGS * SPEED_OF_LIGHT_CGS / PLANCK_CONST_BAR_CGS / frequency;
    }
};
inline double density_to_frequency(double density_CGS) {
    return std::sqrt(4 * PI * ELEMENTARY_CHARGE_CGS * ELEMENTARY_CHARGE_CGS * density_CGS / ELECTRON_MASS_CGS);
}

// ---- species (reference src/Species_register.h, registered at src/Config_reader.cpp:19-22)
struct Species {
    std::string name;
    double charge, mass, charge_to_mass_ratio;
    bool is_photon;
};
class Species_register {
public:
    int register_species(const std::string & name, double charge, double mass, bool is_photon = false) {
        if (find_species(name) >= 0) throw std::invalid_argument("Species [" + name + "] already registered");
        species.push_back(Species{name, charge, mass, mass == 0.0 ? 0.0 : charge / mass, is_photon});
        return static_cast<int>(species.size()) - 1;
    }
    int find_species(const std::string & name) const {
        for (size_t i = 0; i < species.size(); i++) {
            if (species[i].name == name) return static_cast<int>(i);
        }
        return -1;
    }
    const std::vector<Species> & get_species() const { return species; }

private:
    std::vector<Species> species;
};

// ---- parameters
struct Laser_gaussian {
    bool present = false;
    double amplitude = 0;  // a0^2 / 2
    double wx = 1, wy = 1, wz = 1, x0 = 0, y0 = 0, z0 = 0;
};

struct Output_parameters {
    bool output3d = false, output_xy = false, output_bunch = false, output_photons = true;
    double z0 = 0;
};

// beam density profiles, reference src/profiles.cpp:4-45 (same expression order)
inline std::function<double(double, double, double)> gaussian3d(double amplitude, double wx, double wy, double wz, double x0, double y0, double z0) {
    return [=](double x, double y, double z) {
        return amplitude * std::exp(-(x - x0) * (x - x0) / wx / wx - (y - y0) * (y - y0) / wy / wy - (z - z0) * (z - z0) / wz / wz);
    };
}
inline std::function<double(double, double, double)> parabolic3d(double amplitude, double xsize, double rsize, double x0, double y0, double z0) {
    return [=](double x, double y, double z) {
        const double r_squared = (y - y0) * (y - y0) + (z - z0) * (z - z0);
        const double x_rel_squared = (x - x0) * (x - x0);
        if ((r_squared <= rsize * rsize) && (x_rel_squared < xsize * xsize)) {
            return amplitude * (1 - x_rel_squared / xsize / xsize) * (1 - r_squared / rsize / rsize);
        }
        return 0.0;
    };
}
inline std::function<double(double, double, double)> cylindrical3d(double amplitude, double xsize, double rsize, double x0, double y0, double z0) {
    return [=](double x, double y, double z) {
        const double r_squared = (y - y0) * (y - y0) + (z - z0) * (z - z0);
        const double x_rel_squared = (x - x0) * (x - x0);
        return ((r_squared <= rsize * rsize) && (x_rel_squared < xsize * xsize)) ? amplitude : 0.0;
    };
}

// Bunch_parameters of the reference (src/System_parameters.h)
struct Bunch_parameters {
    int ppcx = 1, ppcy = 1, ppcz = 1;
    double gamma = 1;
    std::function<double(double, double, double)> rho;
    int species_id = 0;
};

struct System_parameters {
    double lx = 0, ly = 0, lz = 0, dx = 0, dy = 0, dz = 0, dt = 0, t_end = 0;
    int ppcy = 1, ppcz = 1, magnetic_field_iterations = 1;
    double base_frequency_SI = 1;
    bool qed = false;
    Laser_gaussian laser;
    std::function<double(double, double)> plasma_profile;  // (y, z) -> density, reference default constant2d(-1)
    int bunch_count = 0;
    std::vector<Bunch_parameters> bunch_parameters_array;
    Output_parameters output;
    Species_register species;
};

// plasma channel, reference src/profiles.cpp:53-64
inline std::function<double(double, double)> powerlaw2d(double power, double radius, double y0, double z0, double min_factor, double max_factor,
                                                        double amplitude) {
    return [=](double y, double z) {
        const double r = std::sqrt((y - y0) * (y - y0) + (z - z0) * (z - z0));
        double value = std::pow(r / radius, power);
        if (value > max_factor) value = max_factor;
        if (value < min_factor) value = min_factor;
        return amplitude * value;
    };
}

class Config_reader {
public:
    Config_reader(const std::string & filename, std::ostream & out) : out(out) {
        params.species.register_species("electron", -1, 1);
        params.species.register_species("positron", 1, 1);
        params.species.register_species("proton", 1, 1836.15267343);
        params.species.register_species("photon", 0, 0, true);
        params.plasma_profile = [](double, double) { return -1.0; };

        out << "Reading config " << filename << "\n" << std::endl;
        config = toml_lite::parse_file(filename);

        if (config->contains("base_density_CGS")) {
            if (config->contains("base_frequency_SI")) throw Config_exception("Both [base_density_CGS] and [base_frequency_SI] exist, choose one");
            params.base_frequency_SI = density_to_frequency(required<double>("base_density_CGS", *config));
            std::printf("Set base_frequency_SI = %s\n", fmt_double(params.base_frequency_SI).c_str());
        } else {
            params.base_frequency_SI = optional<double>("base_frequency_SI", 1.0, *config);
        }
        params.lx = required<double>("lx", *config);
        params.ly = required<double>("ly", *config);
        params.lz = required<double>("lz", *config);
        params.dx = required<double>("dx", *config);
        params.dy = required<double>("dy", *config);
        params.dz = required<double>("dz", *config);
        params.dt = optional<double>("dt", params.lx, *config);
        params.t_end = optional<double>("t_end", 0.0, *config);
        params.ppcy = optional<int>("ppcy", 1, *config);
        params.ppcz = optional<int>("ppcz", 1, *config);
        params.magnetic_field_iterations = optional<int>("magnetic_field_iterations", 0, *config);
        params.qed = optional<bool>("qed", false, *config);
        out << std::endl;
        init_laser();
        out << std::endl;
        init_bunch();
        out << std::endl;
        init_plasma_profile();
        out << std::endl;
        init_output();
    }

    const System_parameters & get_parameters() const { return params; }

private:
    template <class T>
    T required(const std::string & name, const toml_lite::Table & t) {
        T v{};
        if (!t.get(name, v)) throw Config_exception("Value [" + name + "] doesn't exist or is of wrong type");
        out << name << " = " << fmt_value(v) << std::endl;
        return v;
    }
    template <class T>
    T optional(const std::string & name, T def, const toml_lite::Table & t) {
        T v{};
        if (t.get(name, v)) {
            out << name << " = " << fmt_value(v) << std::endl;
            return v;
        }
        out << name << " = " << fmt_value(def) << " (default)" << std::endl;
        return def;
    }

    void init_laser() {
        auto it = config->tables.find("laser");
        if (it == config->tables.end()) {
            out << "No [laser] block in config" << std::endl;
            return;
        }
        out << "Parsing [laser] ..." << std::endl;
        const toml_lite::Table & t = *it->second;
        const std::string shape = required<std::string>("shape", t);
        if (shape != "gaussian") throw Config_exception("Laser shape \"" + shape + "\" is not supported, use \"gaussian\".");
        const double a0 = required<double>("a0", t);
        const double xsigma = required<double>("xsigma", t);
        const double ysigma = required<double>("ysigma", t);
        const double zsigma = required<double>("zsigma", t);
        Laser_gaussian & L = params.laser;
        L.present = true;
        L.x0 = optional<double>("x0", xsigma, t);
        L.y0 = optional<double>("y0", 0.5 * params.ly, t);
        L.z0 = optional<double>("z0", 0.5 * params.lz, t);
        // full width at 1/e^2 -> Gauss parameter (reference src/Config_reader.cpp:124-126)
        L.wx = xsigma / std::sqrt(8.0);
        L.wy = ysigma / std::sqrt(8.0);
        L.wz = zsigma / std::sqrt(8.0);
        L.amplitude = 0.5 * a0 * a0;
    }

    // Config_reader::parse_bunch_table (reference src/Config_reader.cpp:138-185): same keys, defaults and echo order
    void echo_bunch_table(const toml_lite::Table & t) {
        Bunch_parameters b;
        b.ppcx = optional<int>("ppcx", 1, t);
        b.ppcy = optional<int>("ppcy", 1, t);
        b.ppcz = optional<int>("ppcz", 1, t);
        b.gamma = required<double>("gamma", t);
        const std::string shape = required<std::string>("shape", t);
        if (shape == "gaussian") {
            const double rho0 = required<double>("rho0", t);
            const double xsigma = required<double>("xsigma", t);
            const double ysigma = required<double>("ysigma", t);
            const double zsigma = required<double>("zsigma", t);
            const double x0 = optional<double>("x0", xsigma, t);
            const double y0 = optional<double>("y0", 0.5 * params.ly, t);
            const double z0 = optional<double>("z0", 0.5 * params.lz, t);
            // full width at 1/e^2 -> Gauss parameter (:157-158)
            b.rho = gaussian3d(rho0, xsigma / std::sqrt(8.0), ysigma / std::sqrt(8.0), zsigma / std::sqrt(8.0), x0, y0, z0);
        } else if (shape == "parabolic" || shape == "cylindrical") {
            const double rho0 = required<double>("rho0", t);
            const double xsize = required<double>("xsize", t);
            const double rsize = required<double>("rsize", t);
            const double x0 = required<double>("x0", t);
            const double y0 = optional<double>("y0", 0.5 * params.ly, t);
            const double z0 = optional<double>("z0", 0.5 * params.lz, t);
            b.rho = shape == "parabolic" ? parabolic3d(rho0, xsize, rsize, x0, y0, z0) : cylindrical3d(rho0, xsize, rsize, x0, y0, z0);
        } else {
            throw Config_exception("Bunch shape \"" + shape + "\" is not supported, use \"gaussian\", \"parabolic\", or \"cylindrical\".");
        }
        const std::string sp = optional<std::string>("species", "electron", t);
        b.species_id = params.species.find_species(sp);
        // the reference stores the -1 of an unknown name and indexes the register with it (undefined behaviour): refuse instead
        if (b.species_id < 0) throw Config_exception("Species \"" + sp + "\" is not registered");
        params.bunch_parameters_array.push_back(b);
        params.bunch_count++;
    }

    void init_bunch() {
        auto t = config->tables.find("bunch");
        auto ta = config->table_arrays.find("bunch");
        if (t != config->tables.end()) {
            out << "Parsing [bunch] ..." << std::endl;
            echo_bunch_table(*t->second);
        } else if (ta != config->table_arrays.end()) {
            out << "Parsing [[bunch]] ..." << std::endl;
            out << "Size " << ta->second.size() << std::endl;
            for (size_t i = 0; i < ta->second.size(); i++) {
                out << "Parsing element " << i << "..." << std::endl;
                echo_bunch_table(*ta->second[i]);
            }
        } else if (config->has_value("bunch")) {
            out << "\"bunch\" cannot be an element, only a table or an array of tables" << std::endl;
        } else {
            out << "No [bunch] or [[bunch]] in config" << std::endl;
        }
    }

    void init_plasma_profile() {
        auto it = config->tables.find("plasma");
        if (it == config->tables.end()) {
            out << "No [plasma] block in config; uniform plasma will be used";
            return;
        }
        out << "Parsing [plasma] ..." << std::endl;
        const toml_lite::Table & t = *it->second;
        const std::string profile = required<std::string>("profile", t);
        if (profile == "powerlaw") {
            const double power = required<double>("power", t);
            const double radius = required<double>("radius", t);
            const double min_factor = optional<double>("min_factor", 0.0, t);
            const double max_factor = optional<double>("max_factor", 1e10, t);
            const double y0 = optional<double>("y0", 0.5 * params.ly, t);
            const double z0 = optional<double>("z0", 0.5 * params.lz, t);
            params.plasma_profile = powerlaw2d(power, radius, y0, z0, min_factor, max_factor, -1.0);
        } else if (profile == "constant") {
            const double rho0 = optional<double>("rho0", -1, t);
            params.plasma_profile = [rho0](double, double) { return rho0; };
        } else {
            throw Config_exception("Plasma profile \"" + profile + "\" is not supported, use \"powerlaw\" or \"constant\".");
        }
    }

    void init_output() {
        auto it = config->tables.find("output");
        if (it == config->tables.end()) {
            out << "No [output] block in config; no output files will be created" << std::endl;
            return;
        }
        out << "Parsing [output] ..." << std::endl;
        const toml_lite::Table & t = *it->second;
        params.output.output3d = optional<bool>("fields_3d", false, t);
        params.output.output_xy = optional<bool>("fields_xy", false, t);
        params.output.output_bunch = optional<bool>("bunch", false, t);
        params.output.z0 = optional<double>("z0", 0.5 * params.lz, t);
    }

    System_parameters params;
    std::shared_ptr<toml_lite::Table> config;
    std::ostream & out;
};

}  // namespace qhost
