// Configuration readers of the mpp:: mirror (product code): the INI dialect of share/ptgs_*.ini
// and the YAML subset of share/*.yaml. See mpp.h.
#include <algorithm>
#include <cctype>
#include <fstream>
#include <sstream>

#include "mpp.h"

namespace mpp
{
namespace
{
std::string trim(const std::string& s)
{
    size_t b = 0, e = s.size();
    while (b < e && std::isspace(static_cast<unsigned char>(s[b]))) b++;
    while (e > b && std::isspace(static_cast<unsigned char>(s[e - 1]))) e--;
    return s.substr(b, e - b);
}
std::string lower(std::string s)
{
    for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    return s;
}
std::string read_file(const std::string& fileName)
{
    std::ifstream f(fileName);
    if (!f.is_open()) throw std::runtime_error("Cannot open file: " + fileName);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
// cut a trailing comment introduced by any of the given markers (outside of quotes/brackets is
// not needed for the files at hand)
std::string strip_comment(const std::string& line, const std::vector<std::string>& markers)
{
    size_t cut = line.size();
    for (const auto& m : markers)
    {
        const size_t p = line.find(m);
        if (p != std::string::npos) cut = std::min(cut, p);
    }
    return line.substr(0, cut);
}
double to_double(const std::string& s, const std::string& what)
{
    try
    {
        size_t       used = 0;
        const double v    = std::stod(s, &used);
        if (trim(s.substr(used)).empty()) return v;
    }
    catch (...)
    {
    }
    throw std::runtime_error("Cannot parse number '" + s + "' for " + what);
}
}  // namespace

// ---- INI -------------------------------------------------------------------------------------------
ConfigFile::ConfigFile(const std::string& fileName) { parse(read_file(fileName)); }
ConfigFile ConfigFile::FromText(const std::string& text)
{
    ConfigFile c;
    c.parse(text);
    return c;
}
void ConfigFile::parse(const std::string& text)
{
    std::map<std::string, std::string> defines;
    std::string                        section;
    std::istringstream                 in(text);
    std::string                        raw;
    while (std::getline(in, raw))
    {
        std::string line = trim(raw);
        if (line.empty() || line[0] == '#' || line[0] == ';') continue;
        if (line.rfind("@define", 0) == 0)
        {
            std::istringstream ls(line.substr(7));
            std::string        name, value;
            ls >> name;
            std::getline(ls, value);
            defines[name] = trim(strip_comment(value, {"//", "#"}));
            continue;
        }
        // ${NAME} substitution
        for (size_t p = line.find("${"); p != std::string::npos; p = line.find("${", p))
        {
            const size_t e = line.find('}', p);
            if (e == std::string::npos) break;
            const std::string name = line.substr(p + 2, e - p - 2);
            const auto        it   = defines.find(name);
            if (it == defines.end()) throw std::runtime_error("Undefined variable ${" + name + "} in config file");
            line.replace(p, e - p + 1, it->second);
            p += it->second.size();
        }
        if (line[0] == '[')
        {
            const size_t e = line.find(']');
            if (e == std::string::npos) throw std::runtime_error("Malformed section header: " + line);
            section = lower(trim(line.substr(1, e - 1)));
            continue;
        }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string key   = lower(trim(line.substr(0, eq)));
        const std::string value = trim(strip_comment(line.substr(eq + 1), {"//", "#", ";"}));
        data_[section][key]     = value;
    }
}
bool ConfigFile::has(const std::string& s, const std::string& k) const
{
    const auto it = data_.find(lower(s));
    return it != data_.end() && it->second.count(lower(k)) != 0;
}
std::string ConfigFile::read_string(const std::string& s, const std::string& k, const std::string& def, bool must) const
{
    if (!has(s, k))
    {
        if (must) throw std::runtime_error("Mandatory config key not found: [" + s + "] " + k);
        return def;
    }
    return data_.at(lower(s)).at(lower(k));
}
double ConfigFile::read_double(const std::string& s, const std::string& k, double def, bool must) const
{
    if (!has(s, k))
    {
        if (must) throw std::runtime_error("Mandatory config key not found: [" + s + "] " + k);
        return def;
    }
    return to_double(read_string(s, k, ""), k);
}
int ConfigFile::read_int(const std::string& s, const std::string& k, int def, bool must) const
{
    return static_cast<int>(read_double(s, k, def, must));
}
std::vector<double> ConfigFile::read_vector(const std::string& s, const std::string& k) const
{
    std::vector<double> out;
    if (!has(s, k)) return out;
    std::string v = read_string(s, k, "");
    for (auto& c : v)
        if (c == '[' || c == ']' || c == ',') c = ' ';
    std::istringstream ls(v);
    std::string        tok;
    while (ls >> tok) out.push_back(to_double(tok, k));
    return out;
}

// ---- YAML subset -----------------------------------------------------------------------------------
Yaml Yaml::FromFile(const std::string& fileName) { return FromText(read_file(fileName)); }
Yaml Yaml::FromText(const std::string& text)
{
    Yaml               root;
    Yaml*              cur = &root;
    std::istringstream in(text);
    std::string        raw;
    while (std::getline(in, raw))
    {
        std::string line = strip_comment(raw, {"#"});
        if (trim(line).empty()) continue;
        const std::string t = trim(line);
        if (t[0] == '%' || t == "---" || t == "...") continue;
        std::string body = t;
        if (t[0] == '-' && (t.size() == 1 || t[1] == ' '))
        {
            // new element of a top-level block sequence
            root.isSeq_ = true;
            root.seq_.emplace_back();
            cur  = &root.seq_.back();
            body = trim(t.substr(1));
            if (body.empty()) continue;
        }
        const size_t colon = body.find(':');
        if (colon == std::string::npos) throw std::runtime_error("YAML: expected 'key: value' in line: " + raw);
        std::string key = trim(body.substr(0, colon)), value = trim(body.substr(colon + 1));
        if (value.size() >= 2 && ((value.front() == '"' && value.back() == '"') || (value.front() == '\'' && value.back() == '\'')))
            value = value.substr(1, value.size() - 2);
        cur->map_[key] = value;
    }
    return root;
}
double Yaml::getDouble(const std::string& key) const
{
    const auto it = map_.find(key);
    if (it == map_.end()) throw std::runtime_error("YAML: missing required key '" + key + "'");
    const std::string v = lower(it->second);
    if (v == "true") return 1.0;
    if (v == "false") return 0.0;
    return to_double(it->second, key);
}
bool Yaml::getBool(const std::string& key) const { return getDouble(key) != 0.0; }
std::vector<double> Yaml::getDoubleVector(const std::string& key) const
{
    const auto it = map_.find(key);
    if (it == map_.end()) throw std::runtime_error("YAML: missing required key '" + key + "'");
    std::string v = it->second;
    if (v.empty() || v.front() != '[' || v.back() != ']') throw std::runtime_error("YAML: key '" + key + "' is not a flow sequence");
    for (auto& c : v)
        if (c == '[' || c == ']' || c == ',') c = ' ';
    std::vector<double> out;
    std::istringstream  ls(v);
    std::string         tok;
    while (ls >> tok) out.push_back(to_double(tok, key));
    return out;
}

}  // namespace mpp
