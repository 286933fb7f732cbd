/*
 * GpuKineticPlugin -- registers the B200 kinetic material with an unmodified Starfish.
 *
 * Add it to the plugin list handed to Starfish.start() (Main.java:53-56 / MainHeadless.java:53-59):
 *     plugins.add(new starfish.core.materials.GpuKineticPlugin());
 * and select it per material in materials.xml with  <material name="O+" type="kinetic_gpu"> ... </material>.
 *
 * Plugins register BEFORE MaterialsModule.init() installs the stock "KINETIC" parser (Starfish.java:149-154,
 * MaterialsModule.java:77-84), which would overwrite a same-named entry; hence the new type name.
 * Precedent: plugins/surface_processing/SurfaceProcessingPlugin.java:19-25.
 */
package starfish.core.materials;

import org.w3c.dom.Element;

import starfish.core.common.Plugin;
import starfish.core.materials.MaterialsModule.MaterialParser;

public class GpuKineticPlugin implements Plugin {
    @Override
    public void register() {
        MaterialsModule.registerMaterialType("KINETIC_GPU", new MaterialParser() {
            @Override
            public Material addMaterial(String name, Element element) {
                return new GpuKineticMaterial(name, element);
            }
        });
    }
}
