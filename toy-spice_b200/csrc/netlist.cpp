// netlist.cpp — host front-end: SPICE deck text -> numbered device table (tsb::Plan).
//
// C++ restatement of the reference's host-side path, which the north star keeps on the host
// and "shared with the reference" (a Go host would call tsb_plan_add_device with what its own
// pkg/netlist + pkg/circuit produced; this file exists because this image has no Go toolchain):
//   pkg/netlist/parser.go:75-158   Parse (title line, '*' truncation, '+' continuation)
//   pkg/netlist/parser.go:160-283  parseLine / parseDotOperator (.model .op .tran .ac .dc)
//   pkg/netlist/parser.go:285-451  parseModel (defaults per model type; `D (` key quirk)
//   pkg/netlist/parser.go:453-725  parseElement / parseVoltageSource / parseCurrentSource
//   pkg/netlist/parser.go:727-748  ParseValue (mantissa * unit multiplier)
//   pkg/netlist/parser.go:752-915  CreateDevice (model -> device parameter resolution)
//   pkg/circuit/circuit.go:48-71   AssignNodeBranchMaps (node / branch numbering)
#include <cstdlib>
#include <cstring>
#include <regex>
#include <sstream>
#include "tsb_internal.hpp"

namespace tsb {
namespace {

struct ParseError { std::string msg; };

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) ++a;
    while (b > a && isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
std::vector<std::string> fields(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }
std::string upper(std::string s) { for (auto& c : s) c = (char)toupper((unsigned char)c); return s; }
bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }
bool ends_with(const std::string& s, const char* p) {
    size_t n = strlen(p);
    return s.size() >= n && s.compare(s.size() - n, n, p) == 0;
}
std::string join(const std::vector<std::string>& v, size_t from) {
    std::string out;
    for (size_t i = from; i < v.size(); ++i) { if (i > from) out += " "; out += v[i]; }
    return out;
}
std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        size_t b = s.find(sep, a);
        if (b == std::string::npos) { out.push_back(s.substr(a)); break; }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    return out;
}
std::string trim_chars(const std::string& s, const char* set) {
    size_t a = 0, b = s.size();
    while (a < b && strchr(set, s[a])) ++a;
    while (b > a && strchr(set, s[b - 1])) --b;
    return s.substr(a, b - a);
}
std::string replace_all(std::string s, const std::string& from, const std::string& to) {
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) { s.replace(pos, from.size(), to); pos += to.size(); }
    return s;
}

double parse_value(const std::string& val) {   // parser.go:727-748
    static const std::regex re(R"(^([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)(meg|[TGMKkmunpf])?s?$)");
    std::smatch m;
    std::string t = trim(val);
    if (!std::regex_match(t, m, re)) throw ParseError{"invalid value format: " + val};
    double num = strtod(m[1].str().c_str(), nullptr);
    if (m[2].matched) {
        const std::string u = m[2].str();
        double mult = 0;
        if (u == "T") mult = 1e12; else if (u == "G") mult = 1e9; else if (u == "meg") mult = 1e6;
        else if (u == "K" || u == "k") mult = 1e3; else if (u == "m") mult = 1e-3; else if (u == "u") mult = 1e-6;
        else if (u == "n") mult = 1e-9; else if (u == "p") mult = 1e-12; else if (u == "f") mult = 1e-15;
        if (mult != 0) num *= mult;      // "M" has no entry in unitMap: value left unscaled
    }
    return num;
}

struct Element {
    std::string type, name;
    std::vector<std::string> nodes;
    double value = 0;
    std::map<std::string, std::string> params;
};
struct Model { std::string type; std::map<std::string, double> params; };

struct Deck {
    std::string title;
    std::vector<Element> elements;
    std::map<std::string, Model> models;
    int analysis = TSB_AN_OP;
    double tstep = 0, tstop = 0, tstart = 0, tmax = 0;
    bool uic = false;
    std::string dc_src;
    double dc_start = 0, dc_stop = 0, dc_inc = 0;
    std::string dc2_src;                       // hardening: second .dc source (nested sweep)
    double dc2_start = 0, dc2_stop = 0, dc2_inc = 0;
    std::string ac_sweep;
    int ac_points = 0;
    double ac_fstart = 0, ac_fstop = 0;
    bool ended = false;                        // hardening: .end seen
};

void parse_model(Deck& d, std::vector<std::string> f) {   // parser.go:285-451
    if (f.size() < 2) throw ParseError{"insufficient model parameters"};
    std::string name = f[0], type_field = f[1], mtype;
    bool has_open = false;
    size_t paren = type_field.find('(');
    if (paren != std::string::npos) {
        mtype = upper(type_field.substr(0, paren));
        has_open = true;
        f.insert(f.begin() + 2, type_field.substr(paren + 1));
    } else mtype = upper(type_field);
    if (mtype != "D" && mtype != "CORE" && mtype != "NPN" && mtype != "PNP" && mtype != "NMOS" && mtype != "PMOS")
        throw ParseError{"unsupported model type: " + mtype};
    std::string param_str;
    if (has_open) {
        std::vector<std::string> pp(f.begin() + 2, f.end());
        if (!pp.empty() && ends_with(pp.back(), ")")) pp.back().pop_back();
        param_str = join(pp, 0);
    } else if (f.size() > 2) {
        param_str = join(f, 2);
        if (ends_with(param_str, ")")) param_str.pop_back();
    }
    size_t star = param_str.find('*');
    if (star != std::string::npos) param_str = param_str.substr(0, star);
    param_str = trim(param_str);

    std::map<std::string, double> p;
    if (mtype == "D") {
        p = {{"is", 1e-14}, {"n", 1.0}, {"rs", 0.0}, {"cj0", 0.0}, {"m", 0.5}, {"vj", 1.0}, {"bv", 100.0},
             {"eg", 1.11}, {"xti", 3.0}, {"tt", 0.0}, {"fc", 0.5}};
    } else if (mtype == "CORE") {
        p = {{"ms", 1.6e6}, {"alpha", 1e-3}, {"a", 1000.0}, {"c", 0.1}, {"k", 2000.0}, {"tc", 1043.0},
             {"beta", 0.0}, {"area", 1e-4}, {"len", 0.1}};
    } else if (mtype == "NPN" || mtype == "PNP") {
        p = {{"is", 1e-16}, {"bf", 100.0}, {"br", 1.0}, {"nf", 1.0}, {"nr", 1.0}, {"vaf", 100.0}, {"var", 100.0},
             {"ikf", 0.01}, {"ikr", 0.01}, {"rc", 0.0}, {"re", 0.0}, {"rb", 0.0}, {"cje", 0.0}, {"vje", 0.75},
             {"mje", 0.33}, {"cjc", 0.0}, {"vjc", 0.75}, {"mjc", 0.33}, {"tf", 0.0}, {"tr", 0.0}, {"xtb", 0.0},
             {"eg", 1.11}, {"xti", 3.0}};
        if (mtype == "PNP") p["type"] = 1.0;
    } else {
        p = {{"level", 1}, {"vto", 0.7}, {"kp", 2e-5}, {"gamma", 0.5}, {"phi", 0.6}, {"lambda", 0.01}, {"rd", 0.0},
             {"rs", 0.0}, {"cbd", 0.0}, {"cbs", 0.0}, {"is", 1e-14}, {"pb", 0.8}, {"cgso", 0.0}, {"cgdo", 0.0},
             {"cgbo", 0.0}, {"cj", 0.0}, {"mj", 0.5}, {"cjsw", 0.0}, {"mjsw", 0.33}, {"tox", 1e-7}, {"l", 10e-6},
             {"w", 10e-6}};
        if (mtype == "PMOS") p["type"] = 1.0;
    }
    for (const std::string& pair : fields(param_str)) {
        std::vector<std::string> parts = split(pair, '=');
        if (parts.size() != 2) continue;
        p[lower(trim(parts[0]))] = parse_value(trim(parts[1]));
    }
    d.models[name] = Model{mtype, p};
}

void parse_dot(Deck& d, const std::string& line) {   // parser.go:176-283
    std::vector<std::string> f = fields(line);
    if (f.empty()) throw ParseError{"invalid analysis command"};
    std::string cmd = lower(f[0]);
    if (cmd == ".model") { parse_model(d, std::vector<std::string>(f.begin() + 1, f.end())); return; }
    if (cmd == ".op") { d.analysis = TSB_AN_OP; return; }
    if (cmd == ".tran") {
        d.analysis = TSB_AN_TRAN;
        if (f.size() < 3) throw ParseError{"insufficient tran parameters, need at least tstep and tstop"};
        d.tstep = parse_value(f[1]);
        d.tstop = parse_value(f[2]);
        for (size_t i = 3; i < f.size(); ++i) {
            if (f[i] == "uic") { d.uic = true; continue; }
            if (i == 3) d.tstart = parse_value(f[i]);
            if (i == 4) d.tmax = parse_value(f[i]);
        }
        if (d.tmax == 0) d.tmax = d.tstep;
        return;
    }
    if (cmd == ".ac") {                       // parser.go:238-261
        d.analysis = TSB_AN_AC;
        if (f.size() < 5) throw ParseError{"insufficient AC parameters, need sweep type, points, fstart, and fstop"};
        d.ac_sweep = upper(f[1]);
        if (d.ac_sweep != "DEC" && d.ac_sweep != "OCT" && d.ac_sweep != "LIN") throw ParseError{"invalid sweep type: " + d.ac_sweep};
        char* endp = nullptr;
        long pts = strtol(f[2].c_str(), &endp, 10);
        if (endp == f[2].c_str() || *endp) throw ParseError{"invalid number of points: " + f[2]};
        d.ac_points = (int)pts;
        d.ac_fstart = parse_value(f[3]); d.ac_fstop = parse_value(f[4]);
        return;
    }
    if (cmd == ".dc") {
        d.analysis = TSB_AN_DC;
        if (f.size() < 5) throw ParseError{"insufficient DC sweep parameters"};
        d.dc_src = f[1];
        d.dc_start = parse_value(f[2]); d.dc_stop = parse_value(f[3]); d.dc_inc = parse_value(f[4]);
        // HARDENING beyond the reference (its parser stops after the first source; cmd/spice/main.go:325 is ready for
        // DCParam.Source2): `.dc src1 start stop inc src2 start stop inc` -> nested sweep, src1 the outer loop
        if (f.size() >= 9) { d.dc2_src = f[5]; d.dc2_start = parse_value(f[6]); d.dc2_stop = parse_value(f[7]); d.dc2_inc = parse_value(f[8]); }
        return;
    }
    throw ParseError{"unsupported analysis type: " + f[0]};
}

Element parse_source(const std::vector<std::string>& f, const char* type) {   // parser.go:563-725
    if (f.size() < 4) throw ParseError{std::string("insufficient ") + (type[0] == 'V' ? "voltage" : "current") + " source parameters"};
    Element e; e.name = f[0]; e.type = type; e.nodes = {f[1], f[2]};
    std::string remaining = replace_all(replace_all(join(f, 3), "(", " ( "), ")", " ) ");
    std::vector<std::string> words = fields(remaining);
    if (words.empty()) throw ParseError{"missing source type"};
    std::string kind = upper(words[0]);
    if (kind == "DC") {
        if (words.size() < 2) throw ParseError{"missing DC value"};
        e.params["type"] = "dc";
        e.value = parse_value(words[1]);
    } else if (kind == "SIN" || kind == "PULSE" || kind == "PWL") {
        std::string k = lower(kind);
        e.params["type"] = k;
        e.params[k] = trim_chars(join(words, 1), "() ");
    } else if (kind == "AC") {
        if (words.size() < 2) throw ParseError{"missing AC magnitude"};
        e.params["type"] = "ac";
        e.value = parse_value(words[1]);
        e.params["phase"] = words.size() > 2 ? words[2] : "0";
    } else throw ParseError{"unsupported source type: " + words[0]};
    return e;
}

Element parse_element(const std::string& line) {   // parser.go:453-561
    std::vector<std::string> f = fields(line);
    if (f.size() < 3) throw ParseError{"invalid element format: " + line};
    Element e; e.name = f[0]; e.type = std::string(1, (char)toupper((unsigned char)f[0][0]));
    const std::string& t = e.type;
    if (t == "V") return parse_source(f, "V");
    if (t == "I") return parse_source(f, "I");
    if (t == "L") {
        e.nodes = {f[1], f[2]};
        for (size_t i = 3; i < f.size(); ++i) {
            std::vector<std::string> pair = split(f[i], '=');
            if (pair.size() == 2) e.params[lower(pair[0])] = pair[1];
            else if (f[i].find('=') == std::string::npos) e.value = parse_value(f[i]);
        }
        return e;
    }
    if (t == "K") {
        if (f.size() < 4) throw ParseError{"insufficient mutual coupling parameters: need coupling name, inductors and coefficient"};
        double k = parse_value(f.back());
        if (k < -1 || k > 1) throw ParseError{"coupling coefficient must be between -1 and 1"};
        if (f.size() - 2 < 2) throw ParseError{"mutual coupling requires at least two inductors"};
        for (size_t i = 1; i + 1 < f.size(); ++i) e.params["ind" + std::to_string(i)] = f[i];
        e.value = k;
        return e;
    }
    if (t == "D") {
        e.nodes = {f[1], f[2]};
        if (f.size() > 3) e.params["model"] = f[3];
        for (size_t i = 4; i < f.size(); ++i) {      // HARDENING (parser.go:530 "TODO: Inline parameters"): D1 a k MODEL Is=.. N=.. Tt=..
            std::vector<std::string> kv = split(f[i], '=');
            if (kv.size() == 2) e.params["inline_" + lower(kv[0])] = kv[1];
        }
        return e;
    }
    if (t == "Q") {
        if (f.size() < 4) throw ParseError{"insufficient BJT parameters: need nodes and model name"};
        e.nodes = {f[1], f[2], f[3]};
        if (f.size() > 4) e.params["model"] = f[4];
        return e;
    }
    if (t == "M") {
        if (f.size() < 6) throw ParseError{"insufficient MOSFET parameters: need nodes and model name"};
        e.nodes = {f[1], f[2], f[3], f[4]};
        e.params["model"] = f[5];
        for (size_t i = 6; i < f.size(); ++i) {
            std::vector<std::string> parts = split(f[i], '=');
            if (parts.size() == 2) e.params[lower(parts[0])] = parts[1];
        }
        return e;
    }
    e.nodes.assign(f.begin() + 1, f.end() - 1);
    e.value = parse_value(f.back());
    return e;
}

void parse_line(Deck& d, const std::string& line_in) {   // parser.go:160-174
    std::string line = std::regex_replace(line_in, std::regex(R"(\s+)"), " ");
    if (d.ended) return;
    {   // HARDENING (parser.go:155 "TODO: .END"): the deck ends here; the reference reports "unsupported analysis type: .end"
        std::string t = lower(line);
        while (!t.empty() && t.back() == ' ') t.pop_back();
        if (t == ".end") { d.ended = true; return; }
    }
    if (starts_with(line, ".")) { parse_dot(d, line); return; }
    d.elements.push_back(parse_element(line));
}

void parse_deck(const std::string& text, Deck& d) {   // parser.go:75-158
    std::vector<std::string> lines = split(text, '\n');
    if (!lines.empty() && lines.back().empty()) lines.pop_back();
    for (auto& ln : lines) if (!ln.empty() && ln.back() == '\r') ln.pop_back();
    if (lines.empty()) return;
    std::string title = lines[0];
    if (starts_with(title, "*")) title = title.substr(1);
    d.title = trim(title);
    std::string current;
    bool continuation = false;
    for (size_t li = 1; li < lines.size(); ++li) {
        const std::string& raw = lines[li];
        std::string line = trim(raw);
        if (line.empty()) {
            if (!current.empty()) { parse_line(d, current); current.clear(); continuation = false; }
            continue;
        }
        size_t star = line.find('*');
        if (star != std::string::npos) {
            line = trim(line.substr(0, star));
            if (line.empty()) continue;
        }
        if (starts_with(line, "+")) {
            line = trim(line.substr(1));
            if (!current.empty()) current += " " + line;
            continuation = true;
            continue;
        }
        if (continuation && starts_with(raw, " ")) {
            if (!current.empty()) current += " " + line;
            continue;
        }
        if (!current.empty()) parse_line(d, current);
        current = line;
        continuation = false;
    }
    if (!current.empty()) parse_line(d, current);
}

double getp(const std::map<std::string, double>& m, const char* k, double dflt) {
    auto it = m.find(k);
    return it == m.end() ? dflt : it->second;
}

void source_params(const Element& e, Dev& dev) {   // parser.go:836-911, 917-1035
    auto it = e.params.find("type");
    std::string t = it == e.params.end() ? "" : it->second;
    if (t == "dc") { dev.ip = {TSB_SRC_DC}; dev.p = {e.value}; return; }
    if (t == "sin") {
        std::vector<std::string> sp = fields(e.params.at("sin"));
        if (sp.size() < 3) throw ParseError{"insufficient SIN parameters"};
        dev.ip = {TSB_SRC_SIN};
        dev.p = {parse_value(sp[0]), parse_value(sp[1]), parse_value(sp[2]), sp.size() > 3 ? parse_value(sp[3]) : 0.0};
        return;
    }
    if (t == "pulse") {
        std::vector<std::string> pp = fields(e.params.at("pulse"));
        if (pp.size() < 7) throw ParseError{"insufficient PULSE parameters"};
        dev.ip = {TSB_SRC_PULSE};
        for (int i = 0; i < 7; ++i) dev.p.push_back(parse_value(pp[i]));
        return;
    }
    if (t == "pwl") {
        std::vector<std::string> pw = fields(e.params.at("pwl"));
        if (pw.size() < 4 || pw.size() % 2 != 0) throw ParseError{"insufficient or invalid PWL parameters, need pairs of time-value"};
        dev.ip = {TSB_SRC_PWL};
        for (auto& s : pw) dev.p.push_back(parse_value(s));
        for (size_t i = 2; i < dev.p.size(); i += 2)
            if (dev.p[i] <= dev.p[i - 2]) throw ParseError{"PWL time points must be strictly increasing"};
        return;
    }
    if (t == "ac") {          // NewACVoltageSource(name, nodes, 0, mag, phase): a DC 0 source in OP/DC/tran
        // magnitude and phase (degrees) ride along behind the DC value: the AC analysis reads them (vsource.go:155-177)
        dev.ip = {TSB_SRC_DC}; dev.p = {0.0, e.value, parse_value(e.params.at("phase"))};
        return;
    }
    throw ParseError{"unsupported source type: " + t};
}

}  // namespace

int plan_from_netlist(const std::string& text, Plan& plan, std::string& err) {
    try {
        Deck d;
        parse_deck(text, d);
        // circuit.go:48-71
        std::map<std::string, int> node_map, branch_map;
        plan.node_names.assign(1, "0");
        for (const Element& e : d.elements)
            for (const std::string& nm : e.nodes) {
                if (nm == "0" || nm == "gnd") continue;
                if (!node_map.count(nm)) { int idx = (int)node_map.size() + 1; node_map[nm] = idx; plan.node_names.push_back(nm); }
            }
        int b = (int)node_map.size() + 1;
        for (const Element& e : d.elements)
            if (e.type == "V" || e.type == "L") branch_map[e.name] = b++;
        plan.n_nodes = (int)node_map.size();
        plan.n_branches = (int)branch_map.size();
        plan.title = d.title;

        std::map<std::string, int> index_of;
        for (const Element& e : d.elements) {
            Dev dev; dev.name = e.name;
            dev.n_nodes = (int)e.nodes.size();
            if (dev.n_nodes > 4) throw ParseError{"too many nodes on element " + e.name};
            for (int i = 0; i < dev.n_nodes; ++i)
                dev.nodes[i] = (e.nodes[i] == "0" || e.nodes[i] == "gnd") ? 0 : node_map[e.nodes[i]];
            auto bi = branch_map.find(e.name);
            dev.branch = bi == branch_map.end() ? 0 : bi->second;
            const std::string& t = e.type;
            auto model_of = [&](const Model*& out) {
                out = nullptr;
                auto mi = e.params.find("model");
                if (mi == e.params.end()) return false;
                auto mm = d.models.find(mi->second);
                if (mm != d.models.end()) out = &mm->second;
                return true;
            };
            if (t == "R") { dev.kind = TSB_R; dev.p = {e.value}; if (dev.n_nodes != 2) throw ParseError{"resistor " + e.name + ": requires exactly 2 nodes"}; }
            else if (t == "C") { dev.kind = TSB_C; dev.p = {e.value}; if (dev.n_nodes != 2) throw ParseError{"capacitor " + e.name + ": requires exactly 2 nodes"}; }
            else if (t == "L") {
                auto ci = e.params.find("core");
                if (ci != e.params.end()) {
                    auto mm = d.models.find(ci->second);
                    if (mm == d.models.end()) throw ParseError{"undefined core model for inductor " + e.name + ": " + ci->second};
                    if (mm->second.type != "CORE") throw ParseError{"invalid core model type for inductor " + e.name + ": " + mm->second.type};
                    int turns = 100;
                    auto ti = e.params.find("turns");
                    if (ti != e.params.end()) {            // strconv.Atoi: whole string must be an integer
                        char* end = nullptr;
                        long v = strtol(ti->second.c_str(), &end, 10);
                        if (end && *end == 0 && !ti->second.empty()) turns = (int)v;
                    }
                    dev.kind = TSB_LCORE;
                    dev.p = {(double)turns, getp(mm->second.params, "area", 1e-4), getp(mm->second.params, "len", 0.1)};
                } else { dev.kind = TSB_L; dev.p = {e.value}; }
            } else if (t == "K") { dev.kind = TSB_K; dev.p = {e.value}; dev.n_nodes = 0; }
            else if (t == "D") {
                if (dev.n_nodes != 2) throw ParseError{"diode " + e.name + ": requires exactly 2 nodes"};
                dev.kind = TSB_D;
                double is = 1e-14, n = 1.0, tt = 0.0;              // diode.go:66-84
                const Model* m; model_of(m);
                if (m) { is = getp(m->params, "is", is); n = getp(m->params, "n", n); tt = getp(m->params, "tt", tt); }
                auto inl = [&](const char* k, double& v) { auto it = e.params.find(std::string("inline_") + k); if (it != e.params.end()) v = parse_value(it->second); };
                inl("is", is); inl("n", n); inl("tt", tt);
                dev.p = {is, n, tt};
            } else if (t == "Q") {
                dev.kind = TSB_Q;
                double ies = 1e-15, ics = 1e-15, af = 0.98, ikf = 1e-3, ikr = 1e-3, vaf = 50.0, var = 50.0;   // bjt.go:87-108
                int pnp = 0;
                const Model* m; model_of(m);
                if (m) {
                    ies = getp(m->params, "ies", ies); ics = getp(m->params, "ics", ics); af = getp(m->params, "alphaf", af);
                    ikf = getp(m->params, "ikf", ikf); ikr = getp(m->params, "ikr", ikr);
                    vaf = getp(m->params, "vaf", vaf); var = getp(m->params, "var", var);
                    auto ty = m->params.find("type");
                    if (ty != m->params.end() && ty->second == 1.0) pnp = 1;
                }
                dev.p = {ies, ics, af, ikf, ikr, vaf, var, 1.0, 1.0};
                dev.ip = {pnp};
            } else if (t == "M") {
                dev.kind = TSB_M;
                const Model* m;
                if (!model_of(m)) throw ParseError{"mosfet " + e.name + ": model not specified"};
                static const char* keys[29] = {"vto", "kp", "gamma", "phi", "lambda", "w", "l", "tox", "cgso", "cgdo",
                    "cgbo", "cbd", "cbs", "cj", "cjsw", "as", "ad", "ps", "pd", "mj", "pb", "uo", "ucrit", "uexp",
                    "vmax", "theta", "eta", "kappa", "delta"};
                static const double dflt[29] = {0.7, 2e-5, 0.5, 0.6, 0.01, 10e-6, 10e-6, 1e-7, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                    0, 0, 0.5, 0.8, 600.0, 1e4, 0, 0, 0, 0, 0.2, 0};      // mosfet.go:144-208
                int level = 1, pmos = 0;
                dev.p.assign(dflt, dflt + 29);
                if (m) {
                    auto lv = m->params.find("level");
                    if (lv != m->params.end()) level = (int)lv->second;
                    auto ty = m->params.find("type");
                    if (ty != m->params.end()) pmos = ty->second == 1.0 ? 1 : 0;
                    for (int i = 0; i < 29; ++i) dev.p[i] = getp(m->params, keys[i], dev.p[i]);
                }
                auto li = e.params.find("l");
                if (li != e.params.end()) { try { dev.p[6] = parse_value(li->second); } catch (ParseError&) {} }
                auto wi = e.params.find("w");
                if (wi != e.params.end()) { try { dev.p[5] = parse_value(wi->second); } catch (ParseError&) {} }
                dev.ip = {level, pmos};
            } else if (t == "V" || t == "I") {
                dev.kind = t == "V" ? TSB_V : TSB_I;
                source_params(e, dev);
            } else throw ParseError{"unsupported device type: " + t};
            index_of[e.name] = (int)plan.devs.size();
            plan.devs.push_back(dev);
        }
        // mutual couplings: resolve inductor names (circuit.go:126-152)
        for (size_t k = 0; k < d.elements.size(); ++k) {
            if (plan.devs[k].kind != TSB_K) continue;
            const Element& e = d.elements[k];
            for (int i = 1;; ++i) {
                auto it = e.params.find("ind" + std::to_string(i));
                if (it == e.params.end()) break;
                auto di = index_of.find(it->second);
                if (di == index_of.end()) throw ParseError{"inductor " + it->second + " not found for mutual coupling " + e.name};
                int kind = plan.devs[di->second].kind;
                if (kind != TSB_L && kind != TSB_LCORE) throw ParseError{"device " + it->second + " is not an inductor component"};
                plan.devs[k].ip.push_back(di->second);
            }
            if (plan.devs[k].ip.size() < 2) throw ParseError{"mutual coupling " + e.name + " requires at least two inductors"};
        }
        plan.analysis = d.analysis;
        plan.tran[0] = d.tstart; plan.tran[1] = d.tstop; plan.tran[2] = d.tstep; plan.tran[3] = d.tmax;
        plan.uic = d.uic ? 1 : 0;
        plan.dc[0] = d.dc_start; plan.dc[1] = d.dc_stop; plan.dc[2] = d.dc_inc;
        plan.dc_src_name = d.dc_src;
        plan.dc_src_dev = -1; plan.dc2_src_dev = -1;
        if (d.analysis == TSB_AN_DC) {
            auto di = index_of.find(d.dc_src);
            if (di == index_of.end() || plan.devs[di->second].kind != TSB_V) throw ParseError{"source " + d.dc_src + " not found"};
            plan.dc_src_dev = di->second;
            if (!d.dc2_src.empty()) {
                auto d2 = index_of.find(d.dc2_src);
                if (d2 == index_of.end() || plan.devs[d2->second].kind != TSB_V) throw ParseError{"source " + d.dc2_src + " not found"};
                plan.dc2_src_dev = d2->second;
                plan.dc2[0] = d.dc2_start; plan.dc2[1] = d.dc2_stop; plan.dc2[2] = d.dc2_inc;
            }
        }
        plan.ac_sweep = d.ac_sweep == "DEC" ? 0 : d.ac_sweep == "OCT" ? 1 : 2;
        plan.ac_points = d.ac_points; plan.ac_f[0] = d.ac_fstart; plan.ac_f[1] = d.ac_fstop;
        return TSB_OK;
    } catch (ParseError& e) {
        err = e.msg;
        return TSB_E_PARSE;
    } catch (std::exception& e) {
        err = e.what();
        return TSB_E_PARSE;
    }
}

}  // namespace tsb
