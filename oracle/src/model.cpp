// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// URDF → Model (see model.h) and Jacobian-based rigid-body quantities via dual numbers.
#include "model.h"

#include <cstring>
#include <functional>
#include <fstream>
#include <map>
#include <sstream>

namespace orc {
namespace {

struct Xml { std::string tag; std::map<std::string, std::string> attr; std::vector<Xml> kids;
  const Xml* first(const std::string& t) const { for (auto& k : kids) if (k.tag == t) return &k; return nullptr; } };

struct XmlParser {
  const std::string& s; size_t i = 0;
  explicit XmlParser(const std::string& str) : s(str) {}
  void skip_ws() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
  bool starts(const char* lit) const { return s.compare(i, strlen(lit), lit) == 0; }
  void skip_misc() {
    for (;;) { skip_ws();
      if (starts("<?")) { i = s.find("?>", i) + 2; }
      else if (starts("<!--")) { i = s.find("-->", i) + 3; }
      else if (starts("<!")) { i = s.find('>', i) + 1; }
      else break; }
  }
  Xml element() {
    skip_misc(); Xml e; if (s[i] != '<') throw std::runtime_error("xml: expected '<'"); ++i;
    size_t st = i; while (!isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') ++i; e.tag = s.substr(st, i - st);
    for (;;) { skip_ws();
      if (s[i] == '/') { i += 2; return e; }
      if (s[i] == '>') { ++i; break; }
      size_t ks = i; while (s[i] != '=' && !isspace((unsigned char)s[i])) ++i; std::string k = s.substr(ks, i - ks);
      while (s[i] != '"' && s[i] != '\'') ++i; char qc = s[i++]; size_t vs = i; while (s[i] != qc) ++i; e.attr[k] = s.substr(vs, i - vs); ++i; }
    for (;;) {
      // text content is ignored
      while (i < s.size() && s[i] != '<') ++i;
      if (starts("</")) { i = s.find('>', i) + 1; return e; }
      if (starts("<!--")) { i = s.find("-->", i) + 3; continue; }
      e.kids.push_back(element());
    }
  }
};

V3<double> parse3(const std::string& str) { std::istringstream is(str); V3<double> v; is >> v.x >> v.y >> v.z; return v; }
M3<double> rpy_to_R(const V3<double>& rpy) { return rot_zyx<double>(rpy.z, rpy.y, rpy.x); }

struct ULink { std::string name; BodyDef in; bool has_inertia = false; };
struct UJoint { std::string name, type, parent, child; M3<double> R = M3<double>::identity(); V3<double> p; int axis = 0; double lo = 0, hi = 0, eff = 0, vel = 0; };

// inertia of a point-mass-free rigid body expressed in another frame: (R, p) maps child coords → parent coords
BodyDef transform_inertia(const BodyDef& b, const M3<double>& R, const V3<double>& p) {
  BodyDef o; o.mass = b.mass; o.com = p + R * b.com; o.I = R * b.I * transpose(R); return o; }
BodyDef add_inertia(const BodyDef& a, const BodyDef& b) {
  if (a.mass == 0.0 && b.mass == 0.0) return a;
  BodyDef o; o.mass = a.mass + b.mass; o.com = (1.0 / o.mass) * (a.mass * a.com + b.mass * b.com);
  auto shift = [&](const BodyDef& x) { V3<double> d = x.com - o.com; M3<double> S = skew(d); M3<double> StS = transpose(S) * S; M3<double> r = x.I; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) += x.mass * StS(i, j); return r; };
  o.I = shift(a) + shift(b); return o; }

}  // namespace

Model load_model(const std::string& urdf_file, const std::vector<double>& default_joint_state, const std::string& ee_frame) {
  std::ifstream f(urdf_file); if (!f) throw std::invalid_argument("URDF file not found: " + urdf_file);
  std::stringstream ss; ss << f.rdbuf(); std::string text = ss.str();
  XmlParser xp(text); Xml robot = xp.element();
  if (robot.tag != "robot") throw std::runtime_error("URDF: root element is not <robot>");
  std::map<std::string, ULink> links; std::map<std::string, UJoint> joints;  // name-keyed (urdfdom ordering)
  for (auto& e : robot.kids) {
    if (e.tag == "link") {
      ULink l; l.name = e.attr.at("name");
      if (const Xml* in = e.first("inertial")) {
        V3<double> xyz, rpy; if (const Xml* o = in->first("origin")) { if (o->attr.count("xyz")) xyz = parse3(o->attr.at("xyz")); if (o->attr.count("rpy")) rpy = parse3(o->attr.at("rpy")); }
        BodyDef b; b.mass = std::stod(in->first("mass")->attr.at("value"));
        const auto& ia = in->first("inertia")->attr; M3<double> I;
        I(0, 0) = std::stod(ia.at("ixx")); I(0, 1) = I(1, 0) = std::stod(ia.at("ixy")); I(0, 2) = I(2, 0) = std::stod(ia.at("ixz"));
        I(1, 1) = std::stod(ia.at("iyy")); I(1, 2) = I(2, 1) = std::stod(ia.at("iyz")); I(2, 2) = std::stod(ia.at("izz"));
        M3<double> Rin = rpy_to_R(rpy); b.I = Rin * I * transpose(Rin); b.com = xyz; l.in = b; l.has_inertia = true;
      }
      links[l.name] = l;
    } else if (e.tag == "joint" && e.attr.count("type")) {
      UJoint j; j.name = e.attr.at("name"); j.type = e.attr.at("type"); j.parent = e.first("parent")->attr.at("link"); j.child = e.first("child")->attr.at("link");
      if (const Xml* o = e.first("origin")) { V3<double> xyz, rpy; if (o->attr.count("xyz")) xyz = parse3(o->attr.at("xyz")); if (o->attr.count("rpy")) rpy = parse3(o->attr.at("rpy")); j.R = rpy_to_R(rpy); j.p = xyz; }
      if (j.type != "fixed") {
        V3<double> ax(1, 0, 0); if (const Xml* a = e.first("axis")) ax = parse3(a->attr.at("xyz"));
        if (ax.x == 1 && ax.y == 0 && ax.z == 0) j.axis = 0; else if (ax.x == 0 && ax.y == 1 && ax.z == 0) j.axis = 1; else if (ax.x == 0 && ax.y == 0 && ax.z == 1) j.axis = 2;
        else throw std::runtime_error("URDF: only +x/+y/+z revolute axes supported: " + j.name);
        if (const Xml* l = e.first("limit")) { auto g = [&](const char* k) { return l->attr.count(k) ? std::stod(l->attr.at(k)) : 0.0; }; j.lo = g("lower"); j.hi = g("upper"); j.eff = g("effort"); j.vel = g("velocity"); }
      }
      joints[j.name] = j;
    }
  }
  // root link
  std::string root; { std::map<std::string, bool> is_child; for (auto& kv : joints) is_child[kv.second.child] = true; for (auto& kv : links) if (!is_child[kv.first]) root = kv.first; }
  Model m; int nj = 0;
  m.body[0] = links[root].has_inertia ? links[root].in : BodyDef();
  m.frames.push_back({root, 0, M3<double>::identity(), V3<double>()});
  // depth-first traversal; children in joint-name order
  struct Ctx { std::string link; int body; M3<double> R; V3<double> p; };
  std::vector<Ctx> stack{{root, 0, M3<double>::identity(), V3<double>()}};
  // recursion via explicit lambda to preserve depth-first/pre-order numbering
  std::function<void(const Ctx&)> visit = [&](const Ctx& c) {
    for (auto& kv : joints) {
      const UJoint& j = kv.second; if (j.parent != c.link) continue;
      M3<double> Rj = c.R * j.R; V3<double> pj = c.p + c.R * j.p;  // joint frame in current body frame
      if (j.type == "fixed") {
        const ULink& cl = links[j.child];
        if (cl.has_inertia) m.body[c.body] = add_inertia(m.body[c.body], transform_inertia(cl.in, Rj, pj));
        m.frames.push_back({j.child, c.body, Rj, pj});
        visit({j.child, c.body, Rj, pj});
      } else {
        if (nj >= NJ) throw std::runtime_error("URDF: more than 18 actuated joints");
        const int jid = nj++; m.joint[jid] = {j.name, c.body, Rj, pj, j.axis, j.lo, j.hi, j.eff, j.vel};
        const ULink& cl = links[j.child]; m.body[jid + 1] = cl.has_inertia ? cl.in : BodyDef();
        m.frames.push_back({j.child, jid + 1, M3<double>::identity(), V3<double>()});
        visit({j.child, jid + 1, M3<double>::identity(), V3<double>()});
      }
    }
  };
  visit(stack[0]);
  if (nj != NJ) throw std::runtime_error("URDF: expected 18 actuated joints");
  const char* feet[4] = {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"};
  for (int i = 0; i < 4; ++i) { m.foot_frame[i] = m.frame_id(feet[i]); if (m.foot_frame[i] < 0) throw std::runtime_error("URDF: missing foot frame"); }
  m.ee_frame = m.frame_id(ee_frame); if (m.ee_frame < 0) throw std::runtime_error("URDF: missing end-effector frame " + ee_frame);
  m.base_frame = m.frame_id("base"); if (m.base_frame < 0) m.base_frame = 0;
  m.mass = total_mass(m);
  // CentroidalModelInfo (createCentroidalModelInfo, SRBD branch) [upstream, recalled]
  for (int i = 0; i < 6; ++i) m.q_nominal[i] = 0.0;
  for (int i = 0; i < NJ; ++i) m.q_nominal[6 + i] = default_joint_state.at(i);
  Kin<double> k; forward_kinematics<double>(m, m.q_nominal, k);
  V3<double> com; for (int b = 0; b < NB; ++b) com = com + m.body[b].mass * body_com(m, k, b); com = (1.0 / m.mass) * com;
  M3<double> Ig;
  for (int b = 0; b < NB; ++b) { M3<double> Iw = k.R[b] * m.body[b].I * transpose(k.R[b]); V3<double> d = body_com(m, k, b) - com; M3<double> S = skew(d); M3<double> StS = transpose(S) * S;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ig(i, j) += Iw(i, j) + m.body[b].mass * StS(i, j); }
  m.I_nominal = Ig; m.com_to_base_nominal = V3<double>(0, 0, 0) - com;
  return m;
}

double total_mass(const Model& m) { double s = 0; for (int b = 0; b < NB; ++b) s += m.body[b].mass; return s; }

namespace {
using D1 = Dual<1, double>;
using DD = Dual<NQ, D1>;
inline void seed(const double* q, const double* v, DD* qd) { for (int k = 0; k < NQ; ++k) { D1 val; val.v = q[k]; val.d[0] = v ? v[k] : 0.0; qd[k] = DD::variable(val, k); } }
// angular-velocity Jacobian column k of a rotation R(q): vee( dR/dq_k R^T ), as a D1 (value, time derivative along v)
inline void ang_col(const M3<DD>& R, int k, D1 out[3]) {
  M3<D1> dR, Rv; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { dR(i, j) = R(i, j).d[k]; Rv(i, j) = R(i, j).v; }
  M3<D1> W = dR * transpose(Rv); out[0] = W(2, 1); out[1] = W(0, 2); out[2] = W(1, 0);
}
}  // namespace

void frame_jacobians(const Model& m, const double* q, const double* v, int frame, Mat& J, Mat& dJ) {
  DD qd[NQ]; seed(q, v, qd); Kin<DD> k; forward_kinematics<DD>(m, qd, k);
  V3<DD> p = frame_pos(m, k, frame); M3<DD> R = frame_rot(m, k, frame);
  J = Mat(6, NQ); dJ = Mat(6, NQ);
  for (int c = 0; c < NQ; ++c) {
    for (int i = 0; i < 3; ++i) { J(i, c) = p[i].d[c].v; dJ(i, c) = p[i].d[c].d[0]; }
    D1 w[3]; ang_col(R, c, w); for (int i = 0; i < 3; ++i) { J(3 + i, c) = w[i].v; dJ(3 + i, c) = w[i].d[0]; }
  }
}

void compute_rbd(const Model& m, const double* q, const double* v, RbdData& o, int what) {
  DD qd[NQ]; seed(q, v, qd); Kin<DD> k; forward_kinematics<DD>(m, qd, k);
  const V3<double> grav(0, 0, -9.81);
  // per-body COM/angular Jacobians and bias accelerations
  std::vector<Mat> Jc(NB, Mat(3, NQ)), Jw(NB, Mat(3, NQ)); std::vector<V3<double>> ac(NB), wd(NB), w(NB), c(NB), cd(NB); std::vector<M3<double>> Iw(NB);
  for (int b = 0; b < NB; ++b) {
    V3<DD> cb = body_com(m, k, b);
    for (int col = 0; col < NQ; ++col) {
      for (int i = 0; i < 3; ++i) { Jc[b](i, col) = cb[i].d[col].v; ac[b][i] += cb[i].d[col].d[0] * v[col]; cd[b][i] += cb[i].d[col].v * v[col]; }
      D1 wc[3]; ang_col(k.R[b], col, wc); for (int i = 0; i < 3; ++i) { Jw[b](i, col) = wc[i].v; wd[b][i] += wc[i].d[0] * v[col]; w[b][i] += wc[i].v * v[col]; }
    }
    for (int i = 0; i < 3; ++i) c[b][i] = cb[i].v.v;
    M3<double> Rb; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rb(i, j) = k.R[b](i, j).v.v;
    Iw[b] = Rb * m.body[b].I * transpose(Rb);
  }
  auto m3mat = [](const M3<double>& A) { Mat r(3, 3); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, j); return r; };
  if (what & 1) {
    o.M = Mat(NQ, NQ); o.nle = Vec(NQ, 0.0);
    for (int b = 0; b < NB; ++b) {
      const double mb = m.body[b].mass;
      o.M = o.M + mb * (Jc[b].T() * Jc[b]) + Jw[b].T() * (m3mat(Iw[b]) * Jw[b]);
      V3<double> f = mb * (ac[b] - grav); V3<double> n = Iw[b] * wd[b] + cross(w[b], Iw[b] * w[b]);
      for (int col = 0; col < NQ; ++col) for (int i = 0; i < 3; ++i) o.nle[col] += Jc[b](i, col) * f[i] + Jw[b](i, col) * n[i];
    }
    for (int i = 0; i < NQ; ++i) for (int j = 0; j < i; ++j) { double s = 0.5 * (o.M(i, j) + o.M(j, i)); o.M(i, j) = o.M(j, i) = s; }
    o.Jfoot = Mat(12, NQ); o.dJfoot = Mat(12, NQ);
    for (int f = 0; f < 4; ++f) {
      V3<DD> p = frame_pos(m, k, m.foot_frame[f]);
      for (int i = 0; i < 3; ++i) { o.foot_pos[f][i] = p[i].v.v; o.foot_vel[f][i] = p[i].v.d[0];
        for (int col = 0; col < NQ; ++col) { o.Jfoot(3 * f + i, col) = p[i].d[col].v; o.dJfoot(3 * f + i, col) = p[i].d[col].d[0]; } }
    }
    frame_jacobians(m, q, v, m.base_frame, o.Jbase, o.dJbase);
    frame_jacobians(m, q, v, m.ee_frame, o.Jee, o.dJee);
  } else {
    for (int f = 0; f < 4; ++f) { V3<DD> p = frame_pos(m, k, m.foot_frame[f]); for (int i = 0; i < 3; ++i) { o.foot_pos[f][i] = p[i].v.v; o.foot_vel[f][i] = p[i].v.d[0]; } }
  }
  { V3<DD> p = frame_pos(m, k, m.ee_frame); M3<DD> R = frame_rot(m, k, m.ee_frame);
    for (int i = 0; i < 3; ++i) { o.ee_pos[i] = p[i].v.v; o.ee_vel[i] = p[i].v.d[0]; for (int j = 0; j < 3; ++j) o.ee_rot(i, j) = R(i, j).v.v; }
    // angular velocity: vee(Rdot R^T)
    M3<double> Rd; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rd(i, j) = R(i, j).v.d[0];
    M3<double> W = Rd * transpose(o.ee_rot); o.ee_angvel = V3<double>(W(2, 1), W(0, 2), W(1, 0)); }
  if (what & 2) {
    V3<double> com; for (int b = 0; b < NB; ++b) com = com + m.body[b].mass * c[b]; com = (1.0 / m.mass) * com; o.com = com;
    o.Ag = Mat(6, NQ); o.dAg_v = Vec(6, 0.0);
    for (int b = 0; b < NB; ++b) {
      const double mb = m.body[b].mass; V3<double> d = c[b] - com; Mat S = m3mat(skew(d));
      o.Ag.add_block(0, 0, Jc[b], mb); o.Ag.add_block(3, 0, m3mat(Iw[b]) * Jw[b]); o.Ag.add_block(3, 0, S * Jc[b], mb);
      V3<double> ang = Iw[b] * wd[b] + cross(w[b], Iw[b] * w[b]) + mb * cross(d, ac[b]);
      for (int i = 0; i < 3; ++i) { o.dAg_v[i] += mb * ac[b][i]; o.dAg_v[3 + i] += ang[i]; }
    }
  }
}

}  // namespace orc
