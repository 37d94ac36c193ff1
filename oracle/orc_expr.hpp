// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// Stand-in for mrpt::expr::CRuntimeCompiledExpression (exprtk) as HolonomicBlend uses it
// (HolonomicBlend.cpp:895-912 symbol table, :931-933 compile, :965-966 eval): arithmetic
// expressions over named double variables bound by pointer, case-insensitive identifiers,
// the constant `pi`, and the usual elementary functions. [MRPT/exprtk, assumed behaviour]
#pragma once
#include <cctype>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc
{
class Expression
{
   public:
    void register_symbol_table(const std::map<std::string, double*>& vars)
    {
        for (const auto& kv : vars) vars_[lower(kv.first)] = kv.second;
    }
    void compile(const std::string& text)
    {
        src_  = text;
        pos_  = 0;
        root_ = parse_sum();
        skip_ws();
        if (pos_ != src_.size()) throw std::runtime_error("expression: trailing characters in '" + src_ + "'");
    }
    bool   compiled() const { return static_cast<bool>(root_); }
    double eval() const
    {
        if (!root_) throw std::runtime_error("expression not compiled");
        return root_->eval();
    }

   private:
    struct Node
    {
        virtual ~Node()             = default;
        virtual double eval() const = 0;
    };
    using P = std::unique_ptr<Node>;
    struct Num : Node
    {
        double v;
        explicit Num(double x) : v(x) {}
        double eval() const override { return v; }
    };
    struct Var : Node
    {
        const double* p;
        explicit Var(const double* q) : p(q) {}
        double eval() const override { return *p; }
    };
    struct Bin : Node
    {
        char op;
        P    a, b;
        Bin(char o, P x, P y) : op(o), a(std::move(x)), b(std::move(y)) {}
        double eval() const override
        {
            const double x = a->eval(), y = b->eval();
            switch (op)
            {
                case '+': return x + y;
                case '-': return x - y;
                case '*': return x * y;
                case '/': return x / y;
                default: return std::pow(x, y);
            }
        }
    };
    struct Neg : Node
    {
        P a;
        explicit Neg(P x) : a(std::move(x)) {}
        double eval() const override { return -a->eval(); }
    };
    struct Call : Node
    {
        std::string    f;
        std::vector<P> args;
        double         eval() const override
        {
            const double x = args[0]->eval();
            if (f == "abs") return std::fabs(x);
            if (f == "sqrt") return std::sqrt(x);
            if (f == "sin") return std::sin(x);
            if (f == "cos") return std::cos(x);
            if (f == "tan") return std::tan(x);
            if (f == "atan") return std::atan(x);
            if (f == "exp") return std::exp(x);
            if (f == "log") return std::log(x);
            if (f == "sgn") return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0);
            const double y = args.at(1)->eval();
            if (f == "min") return std::min(x, y);
            if (f == "max") return std::max(x, y);
            if (f == "atan2") return std::atan2(x, y);
            if (f == "pow") return std::pow(x, y);
            throw std::runtime_error("expression: unknown function " + f);
        }
    };

    static std::string lower(std::string s)
    {
        for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
        return s;
    }
    void skip_ws()
    {
        while (pos_ < src_.size() && std::isspace(static_cast<unsigned char>(src_[pos_]))) pos_++;
    }
    bool eat(char c)
    {
        skip_ws();
        if (pos_ < src_.size() && src_[pos_] == c)
        {
            pos_++;
            return true;
        }
        return false;
    }
    P parse_sum()
    {
        P a = parse_product();
        for (;;)
        {
            if (eat('+'))
                a = P(new Bin('+', std::move(a), parse_product()));
            else if (eat('-'))
                a = P(new Bin('-', std::move(a), parse_product()));
            else
                return a;
        }
    }
    P parse_product()
    {
        P a = parse_unary();
        for (;;)
        {
            if (eat('*'))
                a = P(new Bin('*', std::move(a), parse_unary()));
            else if (eat('/'))
                a = P(new Bin('/', std::move(a), parse_unary()));
            else
                return a;
        }
    }
    P parse_unary()
    {
        if (eat('-')) return P(new Neg(parse_unary()));
        if (eat('+')) return parse_unary();
        P a = parse_primary();
        if (eat('^')) return P(new Bin('^', std::move(a), parse_unary()));
        return a;
    }
    P parse_primary()
    {
        skip_ws();
        if (pos_ >= src_.size()) throw std::runtime_error("expression: unexpected end in '" + src_ + "'");
        const char c = src_[pos_];
        if (c == '(')
        {
            pos_++;
            P a = parse_sum();
            if (!eat(')')) throw std::runtime_error("expression: missing ')' in '" + src_ + "'");
            return a;
        }
        if (std::isdigit(static_cast<unsigned char>(c)) || c == '.')
        {
            size_t       used = 0;
            const double v    = std::stod(src_.substr(pos_), &used);
            pos_ += used;
            return P(new Num(v));
        }
        if (std::isalpha(static_cast<unsigned char>(c)) || c == '_')
        {
            size_t e = pos_;
            while (e < src_.size() && (std::isalnum(static_cast<unsigned char>(src_[e])) || src_[e] == '_')) e++;
            const std::string id = lower(src_.substr(pos_, e - pos_));
            pos_                 = e;
            if (eat('('))
            {
                auto call = new Call();
                P    holder(call);
                call->f = id;
                if (!eat(')'))
                {
                    do
                    {
                        call->args.push_back(parse_sum());
                    } while (eat(','));
                    if (!eat(')')) throw std::runtime_error("expression: missing ')' in call to " + id);
                }
                if (call->args.empty()) throw std::runtime_error("expression: call without arguments: " + id);
                return holder;
            }
            if (id == "pi") return P(new Num(M_PI));
            const auto it = vars_.find(id);
            if (it == vars_.end()) throw std::runtime_error("expression: unknown symbol '" + id + "'");
            return P(new Var(it->second));
        }
        throw std::runtime_error(std::string("expression: unexpected character '") + c + "' in '" + src_ + "'");
    }

    std::map<std::string, double*> vars_;
    std::string                    src_;
    size_t                         pos_ = 0;
    P                              root_;
};

}  // namespace orc
