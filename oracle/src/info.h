// ORACLE — TEST INFRASTRUCTURE ONLY.
// Boost-property-tree INFO reader (stand-in for boost::property_tree::read_info + ocs2::loadData
// used at qm_interface/src/QMInterface.cpp:66-73,147-158,177-231 and qm_wbc/src/WbcBase.cpp:584-594).
#pragma once
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "linalg.h"

namespace orc {

struct InfoNode {
  std::string value;
  std::vector<std::pair<std::string, std::shared_ptr<InfoNode>>> kids;
  const InfoNode* child(const std::string& k) const { for (auto& p : kids) if (p.first == k) return p.second.get(); return nullptr; }
  const InfoNode* path(const std::string& dotted) const {
    const InfoNode* n = this; size_t pos = 0;
    while (n && pos <= dotted.size()) { size_t e = dotted.find('.', pos); std::string k = dotted.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
      n = n->child(k); if (e == std::string::npos) break; pos = e + 1; }
    return n; }
  bool has(const std::string& dotted) const { return path(dotted) != nullptr; }
  double num(const std::string& dotted) const { auto* n = path(dotted); if (!n) throw std::runtime_error("INFO key missing: " + dotted); return std::stod(n->value); }
  double num_or(const std::string& dotted, double dflt) const { auto* n = path(dotted); return n ? std::stod(n->value) : dflt; }
  std::string str(const std::string& dotted) const { auto* n = path(dotted); if (!n) throw std::runtime_error("INFO key missing: " + dotted); return n->value; }
};

inline std::vector<std::string> info_tokenize(const std::string& text) {
  std::vector<std::string> toks; size_t i = 0, n = text.size();
  while (i < n) {
    char ch = text[i];
    if (ch == ';') { while (i < n && text[i] != '\n') ++i; continue; }
    if (ch == '/' && i + 1 < n && text[i + 1] == '/') { while (i < n && text[i] != '\n') ++i; continue; }
    if (ch == '\n') { toks.push_back("\n"); ++i; continue; }
    if (isspace((unsigned char)ch)) { ++i; continue; }
    if (ch == '{' || ch == '}') { toks.push_back(std::string(1, ch)); ++i; continue; }
    if (ch == '"') { size_t e = text.find('"', i + 1); toks.push_back(text.substr(i + 1, e - i - 1)); i = e + 1; continue; }
    size_t s = i; while (i < n && !isspace((unsigned char)text[i]) && text[i] != '{' && text[i] != '}' && text[i] != ';') ++i;
    toks.push_back(text.substr(s, i - s));
  }
  return toks;
}

inline std::shared_ptr<InfoNode> info_parse_file(const std::string& file) {
  std::ifstream f(file); if (!f) throw std::invalid_argument("file not found: " + file);
  std::stringstream ss; ss << f.rdbuf();
  auto toks = info_tokenize(ss.str());
  auto root = std::make_shared<InfoNode>();
  std::vector<InfoNode*> stack{root.get()};
  InfoNode* last = nullptr; size_t i = 0;
  while (i < toks.size()) {
    const std::string& t = toks[i];
    if (t == "\n") { ++i; continue; }
    if (t == "{") { if (!last) throw std::runtime_error("INFO: '{' without key"); stack.push_back(last); last = nullptr; ++i; continue; }
    if (t == "}") { stack.pop_back(); last = nullptr; ++i; continue; }
    auto node = std::make_shared<InfoNode>();
    if (i + 1 < toks.size() && toks[i + 1] != "\n" && toks[i + 1] != "{" && toks[i + 1] != "}") { node->value = toks[i + 1]; i += 2; } else { ++i; }
    stack.back()->kids.emplace_back(t, node); last = node.get();
  }
  return root;
}

// ocs2::loadData::loadEigenMatrix semantics: entries "(i,j) v" times optional "scaling"; missing entries are zero.
inline Mat info_matrix(const InfoNode& root, const std::string& name, int rows, int cols) {
  Mat m(rows, cols); const InfoNode* n = root.path(name); if (!n) throw std::runtime_error("INFO matrix missing: " + name);
  double scaling = n->child("scaling") ? std::stod(n->child("scaling")->value) : 1.0;
  for (auto& kv : n->kids) { int i, j; if (sscanf(kv.first.c_str(), "(%d,%d)", &i, &j) == 2 && i < rows && j < cols) m(i, j) = scaling * std::stod(kv.second->value); }
  return m;
}
inline std::vector<std::string> info_list(const InfoNode& root, const std::string& name) {
  std::vector<std::string> out; const InfoNode* n = root.path(name); if (!n) return out;
  for (auto& kv : n->kids) { int i; if (sscanf(kv.first.c_str(), "[%d]", &i) == 1) { if ((int)out.size() <= i) out.resize(i + 1); out[i] = kv.second->value; } }
  return out;
}

}  // namespace orc
