/*
 * ParityDump -- pins the CPU oracle of this repository against the REAL reference.
 *
 * A Starfish plugin + main(): every golden case of tests/golden/*.npz (exported to plain text by
 * tools/java/export_cases.py) is run through the reference's own public API -- UniformMesh from domain.xml,
 * KineticMaterial from materials.xml, KineticMaterial.addParticle(MeshData, Particle) and
 * KineticMaterial.updateFields() -- and the resulting particle state and fields are written with the raw bits of
 * every double, so tests/test_golden.py::test_oracle_matches_java_reference compares bit for bit.
 *
 * One command on any machine with a JDK (>= 11) and the reference checkout:
 *     tools/java/run_parity_dump.sh /path/to/Starfish
 * then commit tests/golden/java/*.txt: DESIGN.md's "parity unpinned" becomes "pinned by the reference".
 *
 * Input per case (tests/golden/java/<case>/): starfish.xml, domain.xml, materials.xml (standard Starfish input, written
 * by export_cases.py) and case.txt:
 *     steps <n>
 *     bc <RIGHT> <TOP> <LEFT> <BOTTOM>            DomainBoundaryType names
 *     efi <ni*nj hex longs, i*nj+j>   efj <...>
 *     particles <n>  then n lines: x y z u v w mpw  (hex of Double.doubleToRawLongBits)
 */
package starfish.core.materials; // (same package as KineticMaterial: nothing package-private is used, it only keeps the plugin next to its subject)

import java.io.BufferedReader;
import java.io.FileReader;
import java.io.IOException;
import java.io.PrintWriter;
import java.util.ArrayList;
import java.util.Iterator;
import java.util.Locale;
import java.util.StringTokenizer;

import org.w3c.dom.Element;

import starfish.core.common.CommandModule;
import starfish.core.common.Options;
import starfish.core.common.Plugin;
import starfish.core.common.Starfish;
import starfish.core.domain.Field2D;
import starfish.core.domain.Mesh;
import starfish.core.domain.Mesh.DomainBoundaryType;
import starfish.core.domain.Mesh.Face;
import starfish.core.io.InputParser;
import starfish.core.materials.KineticMaterial.MeshData;
import starfish.core.materials.KineticMaterial.Particle;

public class ParityDump implements Plugin {
    static String hex(double d) {
        return Long.toHexString(Double.doubleToRawLongBits(d));
    }

    static double unhex(String s) {
        return Double.longBitsToDouble(Long.parseUnsignedLong(s, 16));
    }

    @Override
    public void register() {
        Starfish.register("parity_dump", new CommandModule() {
            @Override
            public void process(Element element) {
                try {
                    run(InputParser.getValue("case", element), InputParser.getValue("out", element));
                } catch (IOException e) {
                    throw new RuntimeException(e);
                }
            }
        });
    }

    static void run(String caseFile, String outFile) throws IOException {
        Mesh mesh = Starfish.getMeshList().get(0);
        int steps = 0;
        ArrayList<double[]> parts = new ArrayList<>();
        double[] efi = null, efj = null;
        String[] bc = null;
        try (BufferedReader in = new BufferedReader(new FileReader(Starfish.options.wd + caseFile))) {
            String line;
            while ((line = in.readLine()) != null) {
                StringTokenizer t = new StringTokenizer(line);
                if (!t.hasMoreTokens())
                    continue;
                String key = t.nextToken();
                if (key.equals("steps"))
                    steps = Integer.parseInt(t.nextToken());
                else if (key.equals("bc"))
                    bc = new String[] { t.nextToken(), t.nextToken(), t.nextToken(), t.nextToken() };
                else if (key.equals("efi") || key.equals("efj")) {
                    double[] f = new double[mesh.ni * mesh.nj];
                    for (int k = 0; k < f.length; k++)
                        f[k] = unhex(t.nextToken());
                    if (key.equals("efi"))
                        efi = f;
                    else
                        efj = f;
                } else if (key.equals("particles")) {
                    int n = Integer.parseInt(t.nextToken());
                    for (int k = 0; k < n; k++) {
                        StringTokenizer p = new StringTokenizer(in.readLine());
                        double[] v = new double[7];
                        for (int q = 0; q < 7; q++)
                            v[q] = unhex(p.nextToken());
                        parts.add(v);
                    }
                }
            }
        }
        /* mesh faces (Face.val() order RIGHT, TOP, LEFT, BOTTOM), then what <starfish/> would start */
        Face[] faces = { Face.RIGHT, Face.TOP, Face.LEFT, Face.BOTTOM };
        for (int f = 0; f < 4; f++)
            mesh.setMeshBCType(faces[f], DomainBoundaryType.valueOf(bc[f]), 0);
        Starfish.domain_module.start();
        Starfish.materials_module.start(); /* Material.init() */
        for (int f = 0; f < 4; f++) /* (Mesh.init() may reset the faces) */
            mesh.setMeshBCType(faces[f], DomainBoundaryType.valueOf(bc[f]), 0);
        double[][] ei = Starfish.domain_module.getEfi(mesh).getData(), ej = Starfish.domain_module.getEfj(mesh).getData();
        for (int i = 0; i < mesh.ni; i++)
            for (int j = 0; j < mesh.nj; j++) {
                ei[i][j] = efi[i * mesh.nj + j];
                ej[i][j] = efj[i * mesh.nj + j];
            }
        KineticMaterial km = Starfish.getKineticMaterial(0);
        MeshData md = km.getMeshData(mesh);
        for (double[] v : parts) {
            Particle part = new Particle(new double[] { v[0], v[1], v[2] }, new double[] { v[3], v[4], v[5] }, v[6], km);
            part.born_it = 0;
            km.addParticle(md, part); /* KM:759-802 */
        }
        for (int s = 0; s < steps; s++)
            km.updateFields(); /* KM:117-163 */

        try (PrintWriter out = new PrintWriter(Starfish.options.wd + outFile)) {
            out.println("java " + System.getProperty("java.version") + " np " + km.getNp() + " steps " + steps);
            out.println("sums " + hex(km.getMassSum()) + " " + hex(km.getMomentumSum()[0]) + " " + hex(km.getMomentumSum()[1]) + " "
                    + hex(km.getMomentumSum()[2]) + " " + hex(km.getEnergySum()));
            Iterator<Particle> it = km.getIterator(mesh);
            while (it.hasNext()) {
                Particle p = it.next();
                out.println("P " + p.id + " " + hex(p.pos[0]) + " " + hex(p.pos[1]) + " " + hex(p.pos[2]) + " " + hex(p.vel[0]) + " "
                        + hex(p.vel[1]) + " " + hex(p.vel[2]) + " " + hex(p.lc[0]) + " " + hex(p.lc[1]) + " " + hex(p.dt) + " " + hex(p.mpw));
            }
            String[] names = { "nd", "u", "v", "w", "count-sum", "u-sum", "v-sum", "w-sum", "uu-sum", "vv-sum", "ww-sum", "mpc-sum" };
            for (String name : names) {
                Field2D f = km.getFieldManager2d().get(mesh, name);
                StringBuilder sb = new StringBuilder("F " + name);
                for (int i = 0; i < mesh.ni; i++)
                    for (int j = 0; j < mesh.nj; j++)
                        sb.append(' ').append(hex(f.getData()[i][j]));
                out.println(sb);
            }
        }
    }

    /** args: the case directory (holding starfish.xml); runs serial and non-randomised like `-serial -nr` */
    public static void main(String[] args) {
        Locale.setDefault(new Locale("en", "US"));
        ArrayList<Plugin> plugins = new ArrayList<>();
        plugins.add(new ParityDump());
        Options options = new Options(new String[] { "-dir=" + args[0], "-serial", "-nr" });
        new Starfish().start(options, plugins, null);
    }
}
