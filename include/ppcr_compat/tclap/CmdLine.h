// Minimal TCLAP subset: CmdLine, ValueArg<T>, SwitchArg, UnlabeledValueArg<T>, ArgException -- the classes the
// reference CLI uses (src/prob_point_cloud_registration_ex.cc:34-90), same constructor argument order.
#ifndef PPCR_COMPAT_TCLAP_CMDLINE_H
#define PPCR_COMPAT_TCLAP_CMDLINE_H
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace TCLAP {

class ArgException : public std::exception {
public:
    ArgException(const std::string& text, const std::string& id) : text_(text), id_(id) {}
    std::string error() const { return text_; }
    std::string argId() const { return id_; }
    const char* what() const noexcept override { return text_.c_str(); }

private:
    std::string text_, id_;
};

class CmdLine;

class Arg {
public:
    Arg(const std::string& flag, const std::string& name, const std::string& desc, bool req, bool needs_value)
        : flag_(flag), name_(name), desc_(desc), required_(req), needs_value_(needs_value)
    {
    }
    virtual ~Arg() = default;
    bool isSet() const { return set_; }
    const std::string& flag() const { return flag_; }
    const std::string& name() const { return name_; }
    const std::string& description() const { return desc_; }
    bool required() const { return required_; }
    bool needsValue() const { return needs_value_; }
    virtual bool unlabeled() const { return false; }
    virtual void assign(const std::string& text) = 0;
    std::string id() const
    {
        if (unlabeled()) return "Argument: (--" + name_ + ")";
        return flag_.empty() ? "Argument: (--" + name_ + ")" : "Argument: -" + flag_ + " (--" + name_ + ")";
    }

protected:
    std::string flag_, name_, desc_;
    bool required_, needs_value_;
    bool set_ = false;
};

class CmdLine {
public:
    CmdLine(const std::string& message, char delimiter = ' ', const std::string& version = "none")
        : message_(message), version_(version)
    {
        (void)delimiter;
    }
    void add(Arg* a) { args_.push_back(a); }
    void add(Arg& a) { args_.push_back(&a); }
    void parse(int argc, const char* const* argv)
    {
        prog_ = argc > 0 ? argv[0] : "prog";
        std::vector<Arg*> positional;
        for (Arg* a : args_)
            if (a->unlabeled()) positional.push_back(a);
        std::size_t next_pos = 0;
        for (int i = 1; i < argc; ++i) {
            const std::string tok = argv[i];
            if (tok == "-h" || tok == "--help") {
                usage(std::cout);
                std::exit(0);
            }
            if (tok == "--version") {
                std::cout << prog_ << "  version: " << version_ << std::endl;
                std::exit(0);
            }
            Arg* hit = nullptr;
            std::string inline_value;
            bool has_inline = false;
            if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
                std::string name = tok.substr(2);
                const auto eq = name.find('=');
                if (eq != std::string::npos) {
                    inline_value = name.substr(eq + 1);
                    name = name.substr(0, eq);
                    has_inline = true;
                }
                for (Arg* a : args_)
                    if (!a->unlabeled() && a->name() == name) hit = a;
                if (!hit) throw ArgException("Couldn't find match for argument", "Argument: " + tok);
            } else if (tok.size() >= 2 && tok[0] == '-' && !is_number(tok)) {
                const std::string flag = tok.substr(1, 1);
                for (Arg* a : args_)
                    if (!a->unlabeled() && !a->flag().empty() && a->flag() == flag) hit = a;
                if (!hit) throw ArgException("Couldn't find match for argument", "Argument: " + tok);
                if (tok.size() > 2) {
                    inline_value = tok.substr(2);
                    has_inline = true;
                }
            }
            if (hit) {
                if (hit->isSet()) throw ArgException("Argument already set!", hit->id());
                if (hit->needsValue()) {
                    if (has_inline) hit->assign(inline_value);
                    else if (i + 1 < argc) hit->assign(argv[++i]);
                    else throw ArgException("Missing a value for this argument!", hit->id());
                } else {
                    hit->assign("");
                }
            } else {
                if (next_pos >= positional.size()) throw ArgException("Couldn't find match for argument", "Argument: " + tok);
                positional[next_pos++]->assign(tok);
            }
        }
        for (Arg* a : args_)
            if (a->required() && !a->isSet()) throw ArgException("Required argument missing: " + a->name(), a->id());
    }
    void usage(std::ostream& os) const
    {
        os << "\nUSAGE:\n\n   " << prog_;
        for (Arg* a : args_) {
            if (a->unlabeled()) continue;
            os << " [" << (a->flag().empty() ? "--" + a->name() : "-" + a->flag()) << (a->needsValue() ? " <value>" : "") << "]";
        }
        for (Arg* a : args_)
            if (a->unlabeled()) os << " <" << a->name() << ">";
        os << "\n\nWhere:\n\n";
        for (Arg* a : args_) {
            os << "   ";
            if (!a->unlabeled()) {
                if (!a->flag().empty()) os << "-" << a->flag() << ",  ";
                os << "--" << a->name();
            } else {
                os << "<" << a->name() << ">";
            }
            os << "\n     " << (a->required() ? "(required)  " : "") << a->description() << "\n\n";
        }
        os << "   " << message_ << "\n" << std::endl;
    }

private:
    static bool is_number(const std::string& s)
    {
        char* end = nullptr;
        std::strtod(s.c_str(), &end);
        return end && *end == '\0';
    }
    std::string message_, version_, prog_;
    std::vector<Arg*> args_;
};

template <typename T>
class ValueArg : public Arg {
public:
    ValueArg(const std::string& flag, const std::string& name, const std::string& desc, bool req, T value,
             const std::string& type_desc, CmdLine& parser)
        : Arg(flag, name, desc, req, true), value_(value), type_desc_(type_desc)
    {
        parser.add(this);
    }
    T& getValue() { return value_; }
    void assign(const std::string& text) override
    {
        std::istringstream is(text);
        T v;
        if (!(is >> v) || !(is >> std::ws).eof()) throw ArgException("Couldn't read argument value from string '" + text + "'", id());
        value_ = v;
        set_ = true;
    }

protected:
    T value_;
    std::string type_desc_;
};

template <>
inline void ValueArg<std::string>::assign(const std::string& text)
{
    value_ = text;
    set_ = true;
}

template <typename T>
class UnlabeledValueArg : public ValueArg<T> {
public:
    UnlabeledValueArg(const std::string& name, const std::string& desc, bool req, T value, const std::string& type_desc,
                      CmdLine& parser)
        : ValueArg<T>("", name, desc, req, value, type_desc, parser)
    {
    }
    bool unlabeled() const override { return true; }
};

class SwitchArg : public Arg {
public:
    SwitchArg(const std::string& flag, const std::string& name, const std::string& desc, CmdLine& parser, bool def = false)
        : Arg(flag, name, desc, false, false), value_(def), default_(def)
    {
        parser.add(this);
    }
    bool getValue() const { return value_; }
    void assign(const std::string&) override
    {
        value_ = !default_;
        set_ = true;
    }

private:
    bool value_, default_;
};

}  // namespace TCLAP
#endif
