// Runtime math expressions for PTG parameters (product code).
//
// The reference lets the PTG config give |V| and |omega| as text formulas compiled with
// mrpt::expr::CRuntimeCompiledExpression / exprtk (HolonomicBlend.cpp:889-912, 931-933), e.g.
//   expr_W = W_MAX * trimmable_speed * min(1.0, 0.1+abs(dir)/(10*PI/180))
// This is a small shunting-yard compiler to a postfix program over bound double variables:
// + - * / ^, unary minus, parentheses, the constant pi, abs/min/max/sqrt/sin/cos/tan/atan/
// atan2/exp/log/pow/sgn; identifiers are case-insensitive as in exprtk.
#pragma once
#include <cctype>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpp
{
class RuntimeExpr
{
   public:
    void bind(const std::string& name, const double* var) { vars_[lower(name)] = var; }
    /** true if the compiled program reads the bound variable `var` */
    bool references(const double* var) const
    {
        for (const Tok& t : prog_)
            if (t.kind == Tok::VAR && t.var == var) return true;
        return false;
    }

    void compile(const std::string& text)
    {
        prog_.clear();
        std::vector<Tok> ops;  // operator stack
        bool             expectOperand = true;
        size_t           i             = 0;
        auto             fail          = [&](const std::string& m) {
            throw std::runtime_error("expression '" + text + "': " + m);
        };
        auto popToProgram = [&]() {
            prog_.push_back(ops.back());
            ops.pop_back();
        };
        while (i < text.size())
        {
            const char c = text[i];
            if (std::isspace(static_cast<unsigned char>(c)))
            {
                i++;
                continue;
            }
            if (expectOperand)
            {
                if (std::isdigit(static_cast<unsigned char>(c)) || c == '.')
                {
                    size_t used = 0;
                    Tok    t{Tok::NUM};
                    t.value = std::stod(text.substr(i), &used);
                    i += used;
                    prog_.push_back(t);
                    expectOperand = false;
                }
                else if (std::isalpha(static_cast<unsigned char>(c)) || c == '_')
                {
                    size_t e = i;
                    while (e < text.size() && (std::isalnum(static_cast<unsigned char>(text[e])) || text[e] == '_')) e++;
                    const std::string id = lower(text.substr(i, e - i));
                    i                    = e;
                    size_t j             = i;
                    while (j < text.size() && std::isspace(static_cast<unsigned char>(text[j]))) j++;
                    if (j < text.size() && text[j] == '(')
                    {
                        Tok t{Tok::FUNC};
                        t.fn = funcId(id);
                        if (t.fn < 0) fail("unknown function " + id);
                        ops.push_back(t);
                        ops.push_back(Tok{Tok::LPAREN});
                        i = j + 1;
                    }
                    else if (id == "pi")
                    {
                        Tok t{Tok::NUM};
                        t.value = M_PI;
                        prog_.push_back(t);
                        expectOperand = false;
                    }
                    else
                    {
                        const auto it = vars_.find(id);
                        if (it == vars_.end()) fail("unknown symbol " + id);
                        Tok t{Tok::VAR};
                        t.var = it->second;
                        prog_.push_back(t);
                        expectOperand = false;
                    }
                }
                else if (c == '(')
                {
                    ops.push_back(Tok{Tok::LPAREN});
                    i++;
                }
                else if (c == '-')
                {
                    ops.push_back(Tok{Tok::NEG});
                    i++;
                }
                else if (c == '+')
                    i++;
                else
                    fail(std::string("unexpected '") + c + "'");
            }
            else
            {
                if (c == ')')
                {
                    while (!ops.empty() && ops.back().kind != Tok::LPAREN) popToProgram();
                    if (ops.empty()) fail("unbalanced ')'");
                    ops.pop_back();
                    if (!ops.empty() && ops.back().kind == Tok::FUNC) popToProgram();
                    i++;
                }
                else if (c == ',')
                {
                    while (!ops.empty() && ops.back().kind != Tok::LPAREN) popToProgram();
                    if (ops.empty()) fail("',' outside a call");
                    expectOperand = true;
                    i++;
                }
                else if (c == '+' || c == '-' || c == '*' || c == '/' || c == '^')
                {
                    Tok t{Tok::BIN};
                    t.op           = c;
                    const int prec = precedence(t);
                    while (!ops.empty() && ops.back().kind != Tok::LPAREN &&
                           (precedence(ops.back()) > prec || (precedence(ops.back()) == prec && c != '^')))
                        popToProgram();
                    ops.push_back(t);
                    expectOperand = true;
                    i++;
                }
                else
                    fail(std::string("unexpected '") + c + "'");
            }
        }
        if (expectOperand) fail("incomplete expression");
        while (!ops.empty())
        {
            if (ops.back().kind == Tok::LPAREN) fail("unbalanced '('");
            popToProgram();
        }
        eval();  // validates the stack discipline once
    }

    bool compiled() const { return !prog_.empty(); }

    double eval() const
    {
        double st[32];
        int    n = 0;
        for (const Tok& t : prog_)
        {
            switch (t.kind)
            {
                case Tok::NUM: st[n++] = t.value; break;
                case Tok::VAR: st[n++] = *t.var; break;
                case Tok::NEG: st[n - 1] = -st[n - 1]; break;
                case Tok::BIN:
                {
                    const double b = st[--n], a = st[n - 1];
                    st[n - 1] = t.op == '+'   ? a + b
                                : t.op == '-' ? a - b
                                : t.op == '*' ? a * b
                                : t.op == '/' ? a / b
                                              : std::pow(a, b);
                    break;
                }
                case Tok::FUNC:
                    if (t.fn >= kFirstBinaryFn)
                    {
                        const double b = st[--n], a = st[n - 1];
                        st[n - 1]      = call2(t.fn, a, b);
                    }
                    else
                        st[n - 1] = call1(t.fn, st[n - 1]);
                    break;
                default: break;
            }
            if (n <= 0 || n >= 31) throw std::runtime_error("expression: malformed program");
        }
        if (n != 1) throw std::runtime_error("expression: malformed program");
        return st[0];
    }

   private:
    struct Tok
    {
        enum Kind
        {
            NUM,
            VAR,
            BIN,
            NEG,
            FUNC,
            LPAREN
        } kind;
        double        value = 0;
        const double* var   = nullptr;
        char          op    = 0;
        int           fn    = -1;
    };
    static constexpr int kFirstBinaryFn = 100;
    static int           funcId(const std::string& f)
    {
        static const std::map<std::string, int> ids = {
            {"abs", 0}, {"sqrt", 1}, {"sin", 2},   {"cos", 3},   {"tan", 4},     {"atan", 5}, {"exp", 6},
            {"log", 7}, {"sgn", 8},  {"min", 100}, {"max", 101}, {"atan2", 102}, {"pow", 103}};
        const auto it = ids.find(f);
        return it == ids.end() ? -1 : it->second;
    }
    static double call1(int fn, double x)
    {
        switch (fn)
        {
            case 0: return std::fabs(x);
            case 1: return std::sqrt(x);
            case 2: return std::sin(x);
            case 3: return std::cos(x);
            case 4: return std::tan(x);
            case 5: return std::atan(x);
            case 6: return std::exp(x);
            case 7: return std::log(x);
            default: return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0);
        }
    }
    static double call2(int fn, double a, double b)
    {
        switch (fn)
        {
            case 100: return std::min(a, b);
            case 101: return std::max(a, b);
            case 102: return std::atan2(a, b);
            default: return std::pow(a, b);
        }
    }
    static int precedence(const Tok& t)
    {
        if (t.kind == Tok::NEG) return 3;
        if (t.kind == Tok::FUNC) return 5;
        if (t.op == '^') return 4;
        return (t.op == '*' || t.op == '/') ? 2 : 1;
    }
    static std::string lower(std::string s)
    {
        for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
        return s;
    }

    std::map<std::string, const double*> vars_;
    std::vector<Tok>                     prog_;
};

}  // namespace mpp
